"""``MCTS``: the reference's search object API (MCTS.py:15-148) over the CUDA tree kernel.

    mcts = MCTS(nnet, args)              # args.numMCTSSims, args.cpuct
    pi = mcts.getActionProb(game, temp)  # float64[343] (int8 one-hot at temp == 0)

One object owns one device tree (a ``BatchedMCTS`` with a single game); ``getActionProb`` runs the
``numMCTSSims`` simulations as kernel waves, keeps the tables between calls exactly as the reference does
(Coach.py:110 builds one MCTS per episode) and reads the root visit counts back.  A device net is called through
``nnet.predict_head`` (raw float16 head rows; softmax / tanh in the tree kernel) or ``nnet.predict_batch``, any other
net through the reference-style ``nnet.predict(planes)`` once per leaf.

Random draws (leaf symmetry, MCTS.py:92; arg-max tie break, MCTS.py:53) come from the device's counter-based
Philox stream, keyed by a seed taken from numpy's global RNG at construction -- so ``np.random.seed`` still
makes a run reproducible, but the draws are not numpy's MT19937 stream.  For thousands of
concurrent games use ``BatchedMCTS`` / ``SelfPlayEngine`` directly.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .Game import Game
from .backend import default_backend
from .mcts_batched import BatchedMCTS

EPS = 1e-8
log = logging.getLogger(__name__)


class MCTS:
    def __init__(self, nnet, args, seed: int | None = None, backend=None, game_id: int = 0):
        if not hasattr(nnet, "predict"):
            raise TypeError("nnet must provide predict(planes) (NeuralNet.py:27-38)")
        self.nnet = nnet
        self.args = args
        self._backend = backend
        # Random draws (leaf symmetries, arg-max tie-break) are Philox values keyed by (seed; game_id, move, simulation).
        # Without a seed every object draws its own from numpy's global stream, like the reference's use of np.random;
        # with one, the searches are reproducible and ``reset()`` hands out consecutive game ids (Arena resets the
        # player once per game: game i is searched on stream i, as BatchedArena does).
        self._seeded = seed is not None
        self._seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._game_id, self._resets = int(game_id), 0
        self._tree: BatchedMCTS | None = None
        self._last_turn = -1

    # ---- net adapter ----------------------------------------------------------------------------------
    def _evaluate(self, nn_in: torch.Tensor, out: torch.Tensor | None = None):
        if out is not None:                                               # fused head: raw binary16 rows for the tree kernel
            return self.nnet.predict_head(nn_in, out)
        if hasattr(self.nnet, "predict_batch") and getattr(self.nnet, "device", nn_in.device) == nn_in.device:
            return self.nnet.predict_batch(nn_in)
        pis, vs = [], []
        for row in nn_in.cpu().numpy().astype(np.int64):                  # what the reference hands to predict
            pi, v = self.nnet.predict(row)
            pis.append(np.asarray(pi, dtype=np.float32).reshape(-1))
            vs.append(np.asarray(v, dtype=np.float32).reshape(-1)[0])
        return (torch.as_tensor(np.stack(pis)).to(nn_in.device), torch.as_tensor(np.asarray(vs, dtype=np.float32)).to(nn_in.device))

    def _run_search(self) -> None:
        """The numMCTSSims loop.  With a device net (``predict_batch`` on the tree's GPU, and a ``weights_version`` to
        tell when it changes) every (net, wave) step is replayed from a CUDA graph -- identical results, one launch per
        simulation instead of a dozen; ``args.captureSearch = False`` keeps the eager loop.  Host nets (the
        reference-style ``predict``) are called leaf by leaf, exactly as often as the reference calls them."""
        t, n = self._tree, self.nnet
        if (t.be.is_cuda and dict(self.args).get("captureSearch", True) and hasattr(n, "predict_batch")
                and hasattr(n, "weights_version") and getattr(n, "device", None) == t.be.device):
            t.search_captured(n.predict_head if t.fused_head else n.predict_batch, version=(id(n), n.weights_version))
        else:
            t.search_adaptive(self._evaluate)
        t.raise_on_faults()

    def _install_root(self, game) -> None:
        sims, cpuct = int(self.args.numMCTSSims), self.args.cpuct
        state = game._state.copy()
        state["tag"] = self._game_id
        if self._tree is None:
            # a float16 device net on the tree's GPU hands its raw head output (343 logits + value pre-activation) to the
            # tree kernel, which applies softmax / tanh itself -- as the batched self-play does (Coach._evaluator): six
            # small launches fewer per simulation.  args.fusedHead = False keeps predict_batch's float32 (pi, v).
            be, n = self._backend or default_backend(), self.nnet
            fused = bool(be.is_cuda and dict(self.args).get("fusedHead", True) and hasattr(n, "predict_head")
                         and getattr(n, "device", None) == be.device and getattr(n, "infer_dtype", None) == torch.float16)
            self._tree = BatchedMCTS(1, sims, cpuct, seed=self._seed, backend=be, fused_head=fused,
                                     nodes_per_game=min(16383, max(4096, 16 * sims)),
                                     terminal_sims_per_wave=0,   # one game: fewest launches, no tail to bound
                                     # args.leavesPerWave (not in the reference; default 1 = its sequential search): K > 1
                                     # searches leaf-parallel with virtual loss -- batch-K net calls instead of batch-1
                                     leaves_per_wave=int(dict(self.args).get("leavesPerWave", 1)))
            self._tree.reset(state)
        else:
            self._tree.set_params(sims, cpuct)
            if game._curr_turn < self._last_turn:
                # nodes at or below the previous root's turn were reclaimed; the reference would still hold them
                log.warning("MCTS: root turn moved backwards (%d -> %d); reusing a pruned tree",
                            self._last_turn, game._curr_turn)
            self._tree.set_roots(state)
        self._last_turn = game._curr_turn

    # ---- MCTS.py:33-60 ----------------------------------------------------------------------------------
    def getActionProb(self, game, temp=1) -> np.ndarray:
        if not isinstance(game, Game):
            raise TypeError("game must be a Game")
        if not (isinstance(temp, (int, float)) and temp >= 0):
            raise ValueError("temp must be a non-negative number")
        self._install_root(game)
        self._run_search()
        counts = self._tree.root_stats()["counts"][0].cpu().numpy().astype(np.int64)
        if temp == 0:
            best = int(self._tree.sample_actions(0)[0])                   # uniform among the arg-max ties, on device
            probs = np.zeros(len(counts), dtype=np.int8)
            probs[best] = 1
            return probs
        counts = counts ** (1.0 / temp)
        return counts / counts.sum()

    def search(self, game):
        """One simulation from ``game`` (MCTS.py:62-145).  Callers of the reference only reach this through
        ``getActionProb``; provided for scripts that step the search by hand.  Returns None."""
        saved = self.args.numMCTSSims
        try:
            self.args["numMCTSSims"] = 1
            self._install_root(game)
            self._run_search()
        finally:
            self.args["numMCTSSims"] = saved

    def reset(self):
        if not self._seeded:
            return MCTS(self.nnet, self.args, backend=self._backend)      # MCTS.py:147-148
        self._resets += 1
        return MCTS(self.nnet, self.args, seed=self._seed, backend=self._backend, game_id=self._game_id + self._resets - 1)

    def dump(self) -> dict:
        """Root tables of the last search (the reference's Qsa/Nsa/Ns for the root state)."""
        out = self._tree.root_stats(want_q=True)
        return {k: v[0].cpu().numpy() for k, v in out.items()}
