"""inflexion_b200 -- B200-native self-play hot path (batched MCTS + Inflexion rules) behind the reference's API.

Import as ``inflexion_b200`` (see ``inflexion_b200.py`` at the repo root).  The CUDA library
(``libifx_b200.so``, built by ``__graft_entry__.build()``) is loaded on first use and every compute entry
point raises if it, or a CUDA device, is missing: there is no CPU fallback.
"""
__version__ = "0.1.0"
