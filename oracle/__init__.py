"""CPU oracle (test infrastructure only): numpy restatement of the reference's rules + MCTS hot path.

Importable only from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm.
"""
