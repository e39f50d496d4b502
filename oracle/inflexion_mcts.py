"""CPU oracle for the MCTS half of the hot path.  TEST INFRASTRUCTURE ONLY.

Restates ``MCTS.py`` (``getActionProb`` :33-60, ``search`` :62-145) and the sampling step of
``Coach.executeEpisode`` (``Coach.py:66-82``) of Meatssauce/alpha-zero-general-inflexion with every
floating-point operation written out in the precision numpy 2.x (NEP 50) actually uses, so the CUDA tree
kernel can be compared bit for bit.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.

Parity pin: no reference test pins this path ("parity unpinned" by the reference's own tests, SURVEY.md
section 4 / 8c).  The pin is the live reference: ``tests/golden/make_golden.py`` runs the reference's own
``MCTS`` + ``InflexionGame`` (imported from /root/reference in the build container, numpy 2.3.5) with a
deterministic net stub and injected random draws and freezes root visit counts, chosen actions and Q bit
patterns under ``tests/golden/``; ``tests/test_oracle_mcts.py`` checks this oracle against those files.

The reference draws from numpy's global MT19937 stream; a device cannot replay that stream, so every
random draw is INJECTED through a ``Draws`` object.  ``PhiloxDraws`` is the counter-based
generator the CUDA kernels implement (Philox4x32-10 keyed by seed, counter = game/move/sim/purpose).
"""
from __future__ import annotations

import math

import numpy as np

from .inflexion_rules import ACTIONS, ONGOING, OUTCOME_VALUE, OracleGame, random_symmetry

f32 = np.float32
EPS = 1e-8                      # MCTS.py:11

# ---------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) -- the counter-based RNG shared with the device code
# ---------------------------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF

PURPOSE_SYMMETRY = 0            # InflexionGame.py:120-121 (three draws per expansion)
PURPOSE_TIEBREAK = 1            # MCTS.py:53
PURPOSE_ACTION = 2              # Coach.py:81


def philox4x32(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class PhiloxDraws:
    """Random draws keyed by (seed, game, move, sim, purpose); identical to csrc/ifx_philox.cuh."""

    def __init__(self, seed: int, game: int = 0):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.game = game

    def _words(self, move, sim, purpose):
        return philox4x32((self.game & _MASK, move & _MASK, sim & _MASK, purpose), self.key)

    def symmetry(self, move: int, sim: int):
        """(k, shift, axis) for random_symmetry of the leaf expanded by simulation ``sim`` of ``move``."""
        w = self._words(move, sim, PURPOSE_SYMMETRY)
        return w[0] % 6, w[1] % 7, w[2] % 3

    def tiebreak(self, move: int, n: int) -> int:
        return self._words(move, 0, PURPOSE_TIEBREAK)[0] % n

    def uniform(self, move: int) -> float:
        """53-bit uniform in [0,1), built like numpy's random_sample: (a>>5, b>>6)."""
        w = self._words(move, 0, PURPOSE_ACTION)
        return ((w[0] >> 5) * 67108864.0 + (w[1] >> 6)) / 9007199254740992.0


# ---------------------------------------------------------------------------------------------
# numpy's float32 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_FLOAT)
# written out so the test of ``ndarray.sum`` order does not depend on the numpy build
# ---------------------------------------------------------------------------------------------

def pairwise_sum_f32(a: np.ndarray) -> np.float32:
    n = len(a)
    if n < 8:
        res = f32(0.0)
        for x in a:
            res = f32(res + x)
        return res
    if n <= 128:
        r = [f32(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f32(r[j] + a[i + j])
            i += 8
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        while i < n:
            res = f32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f32(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


# ---------------------------------------------------------------------------------------------
# Values carry their numpy "kind": net values are float32, terminal values are Python numbers
# (weak scalars -> float64 arithmetic among themselves).  SURVEY.md appendix B.3.
# ---------------------------------------------------------------------------------------------

class Val:
    __slots__ = ("x", "is32")

    def __init__(self, x, is32: bool):
        self.x = f32(x) if is32 else float(x)
        self.is32 = is32

    def neg(self) -> "Val":
        return Val(-self.x, self.is32)

    def as_f32(self) -> np.float32:
        return f32(self.x)


class OracleMCTS:
    """Dict-keyed PUCT search with the reference's exact table semantics.

    predict(planes int64[4,7,7]) -> (policy float32[343], value float32 scalar)   (NNet.py:78-94)
    """

    def __init__(self, predict, num_sims: int, cpuct, draws, exact_sum: bool = True):
        self.predict = predict
        self.num_sims = num_sims                      # args.numMCTSSims, MCTS.py:45
        self.cpuct = cpuct                            # args.cpuct, MCTS.py:122
        self.draws = draws
        self.exact_sum = exact_sum                    # True: hand-written pairwise sum; False: ndarray.sum()
        self.P, self.V, self.Ns = {}, {}, {}          # MCTS.py:25-31 (Es is never used there)
        self.N, self.Q = {}, {}
        self.move = 0                                 # root-move counter (RNG coordinate)
        self.sim = 0
        self.stats = {"searches": 0, "visits": 0, "expansions": 0, "terminals": 0}

    # -- MCTS.py:62-145 ------------------------------------------------------------------------
    def search(self, game: OracleGame) -> Val:
        s = game.key()
        self.stats["searches"] += 1
        if game.outcome != ONGOING:                   # MCTS.py:85-87
            self.stats["terminals"] += 1
            return Val(-OUTCOME_VALUE[game.outcome], False)
        if s not in self.P:                           # MCTS.py:89-112
            k, t, ax = self.draws.symmetry(self.move, self.sim)
            policy, v = self.predict(random_symmetry(game.to_planes(), k, t, ax))
            valid = game.valid_actions_mask()
            policy = (policy.astype(f32) * valid.astype(f32)).astype(f32)
            total = pairwise_sum_f32(policy) if self.exact_sum else policy.sum()
            if total > 0:
                policy = (policy / f32(total)).astype(f32)
            else:                                     # MCTS.py:100-107
                policy = valid.astype(f32)
                policy = (policy / f32(pairwise_sum_f32(policy))).astype(f32)
            self.P[s], self.V[s], self.Ns[s] = policy, valid, 0
            self.stats["expansions"] += 1
            return Val(-f32(v), True)
        self.stats["visits"] += 1
        valid, pol, ns = self.V[s], self.P[s], self.Ns[s]
        best_u, best_a = -math.inf, -1
        c32 = f32(self.cpuct)
        root_seen = f32(math.sqrt(ns))                # weak Python float -> float32 (B.1)
        root_unseen = f32(math.sqrt(ns + EPS))
        for a in np.nonzero(valid)[0]:                # ascending order, strict '>' (MCTS.py:119-129)
            a = int(a)
            cp = f32(c32 * pol[a])
            if (s, a) in self.Q:
                u = f32(self.Q[(s, a)].as_f32() + f32(f32(cp * root_seen) / f32(1 + self.N[(s, a)])))
            else:
                u = f32(cp * root_unseen)
            if u > best_u:
                best_u, best_a = u, a
        if best_a < 0:
            raise RuntimeError("no valid action (the reference would crash here, MCTS.py:116,131)")
        a = best_a
        v = self.search(game.to_next_state(a))        # MCTS.py:132-134
        if (s, a) in self.Q:                          # MCTS.py:136-142 with NEP-50 promotion (B.3)
            n, q = self.N[(s, a)], self.Q[(s, a)]
            if not q.is32 and not v.is32:
                newq = Val((n * q.x + v.x) / (n + 1), False)
            elif q.is32:
                newq = Val(f32(f32(f32(n) * q.x) + f32(v.x)) / f32(n + 1), True)
            else:
                newq = Val(f32(f32(n * q.x) + v.x) / f32(n + 1), True)
            self.Q[(s, a)], self.N[(s, a)] = newq, n + 1
        else:
            self.Q[(s, a)], self.N[(s, a)] = v, 1
        self.Ns[s] += 1
        return v.neg()

    # -- MCTS.py:33-60 --------------------------------------------------------------------------
    def root_counts(self, game: OracleGame) -> np.ndarray:
        s = game.key()
        return np.array([self.N.get((s, a), 0) for a in range(ACTIONS)], dtype=np.int64)

    def get_action_prob(self, game: OracleGame, temp=1) -> np.ndarray:
        for i in range(self.num_sims):
            self.sim = i
            self.search(game)
        counts = self.root_counts(game)
        if temp == 0:
            best = np.nonzero(counts == counts.max())[0]
            probs = np.zeros(ACTIONS, dtype=np.int8)
            probs[best[self.draws.tiebreak(self.move, len(best))]] = 1
            return probs
        counts = counts ** (1.0 / temp)
        return counts / counts.sum()


def sample_action(pi: np.ndarray, u: float) -> int:
    """np.random.choice(343, p=pi) with the uniform injected (Coach.py:81)."""
    cdf = np.cumsum(pi.astype(np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def self_play_moves(predict, num_sims, cpuct, draws, temp_threshold=30, max_moves=None, max_turns=343,
                    exact_sum=True):
    """The search/sample/advance loop of Coach.executeEpisode (Coach.py:66-88), one record per move.

    Yields dicts with the root counts, pi, the chosen action and the game after the move.
    """
    game = OracleGame(max_turns=max_turns)
    mcts = OracleMCTS(predict, num_sims, cpuct, draws, exact_sum=exact_sum)
    step = 0
    while True:
        step += 1
        temp = int(step < temp_threshold)             # Coach.py:66
        mcts.move = step - 1
        pi = mcts.get_action_prob(game, temp=temp)
        counts = mcts.root_counts(game)
        action = sample_action(pi, draws.uniform(step - 1))
        planes = game.to_planes()
        game = game.to_next_state(action)
        yield {"move": step - 1, "counts": counts, "pi": pi, "action": action, "planes": planes,
               "game": game, "mcts": mcts}
        if game.outcome != ONGOING or (max_moves is not None and step >= max_moves):
            return
