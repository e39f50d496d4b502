"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference mounted, numpy 2.x, torch CPU):

    python tests/golden/make_golden.py

It imports the reference's own ``InflexionGame``, ``MCTS`` and ``Coach`` unmodified, drives them with the
deterministic stub net (``inflexion_b200.stubnet``) and with numpy's global RNG call sites (SURVEY.md
appendix E) replaced by the Philox draws the device uses, and freezes what it observes.  The fixtures are
what the oracle (``oracle/``) and the CUDA path are checked against on machines that do not have the
reference (the GPU box).  Nothing here is imported by the product.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("IFX_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import inflexion_b200  # noqa: E402,F401  (registers the package alias)
from inflexion_b200.stubnet import stub_policy_value  # noqa: E402
from oracle.inflexion_mcts import PhiloxDraws, sample_action  # noqa: E402

from flags import GameOutcome, PlayerColour  # noqa: E402  (reference)
from inflexion.InflexionGame import InflexionGame  # noqa: E402  (reference)
from inflexion.pytorch.NNet import NNetWrapper  # noqa: E402  (reference)
from MCTS import MCTS  # noqa: E402  (reference)
from Coach import Coach  # noqa: E402  (reference)
from utils import dotdict  # noqa: E402  (reference)

OUTCOME_CODE = {GameOutcome.ONGOING: 0, GameOutcome.DRAW: 1, GameOutcome.WON: 2, GameOutcome.LOST: 3}


# ------------------------------------------------------------------------------------------------
# rules transitions
# ------------------------------------------------------------------------------------------------

def play_random_games(rng, n_games, max_turns, rows):
    for _ in range(n_games):
        g = InflexionGame(7, max_turns=max_turns, max_power=6)
        while g.outcome == GameOutcome.ONGOING:
            mask = g.valid_actions_mask()
            acts = np.nonzero(mask)[0]
            # bias toward spreads now and then so wipe-outs and 6->0 overflows show up
            spreads = acts[acts < 294]
            a = int(rng.choice(spreads)) if len(spreads) and rng.random() < 0.35 else int(rng.choice(acts))
            nxt = g.to_next_state(a)
            rows.append(dict(
                board=g._board.reshape(49).astype(np.int8), player=g.player.num, turn=g._curr_turn,
                max_turns=max_turns, action=a, mask=np.packbits(mask.astype(np.uint8)),
                planes=g.to_planes().reshape(4, 49).astype(np.int16), score=g.score(),
                next_board=nxt._board.reshape(49).astype(np.int8), next_outcome=OUTCOME_CODE[nxt.outcome],
                next_player=nxt.player.num, next_turn=nxt._curr_turn))
            g = nxt


def crafted_positions(rows):
    """Hand-built boards for the corners random play rarely reaches."""
    R, B = PlayerColour.RED, PlayerColour.BLUE

    def add(board, player, turn, max_turns, actions=None):
        g = InflexionGame(7, curr_player=player, board=np.array(board, dtype=int).reshape(7, 7),
                          curr_turn=turn, max_turns=max_turns)
        mask = g.valid_actions_mask()
        for a in (np.nonzero(mask)[0] if actions is None else actions):
            a = int(a)
            nxt = g.to_next_state(a)
            rows.append(dict(
                board=g._board.reshape(49).astype(np.int8), player=g.player.num, turn=g._curr_turn,
                max_turns=max_turns, action=a, mask=np.packbits(mask.astype(np.uint8)),
                planes=g.to_planes().reshape(4, 49).astype(np.int16), score=g.score(),
                next_board=nxt._board.reshape(49).astype(np.int8), next_outcome=OUTCOME_CODE[nxt.outcome],
                next_player=nxt.player.num, next_turn=nxt._curr_turn))

    z = np.zeros(49, dtype=int)
    # wipe-out by spread (mover WON), incl. the case where the mover's own stack also vanishes
    b = z.copy(); b[0] = 6; b[7] = -1; add(b, R, 10, 343)
    b = z.copy(); b[0] = 1; b[7] = -6; add(b, R, 10, 343)          # 6+1 -> 0: both sides empty, still WON
    b = z.copy(); b[24] = -3; b[31] = 2; b[38] = 6; add(b, B, 11, 343)
    # spawn forbidden: total power 49 (> 48), and exactly 48 (allowed)
    b = z.copy(); b[:8] = 6; b[8] = 1; b[20] = -1; add(b, R, 50, 343)
    b = z.copy(); b[:8] = 6; b[20] = -1; add(b, R, 50, 343)
    b = z.copy(); b[:8] = -6; b[8] = -1; b[20] = 1; add(b, B, 51, 343)
    # max-turn adjudication: diff >= 2, <= -2 and the draw band, for both colours
    for diff_board in ([3, -1], [1, -3], [2, -1], [1, -1], [1, -2]):
        b = z.copy(); b[3] = diff_board[0]; b[40] = diff_board[1]
        add(b, R, 343, 343); add(b, B, 343, 343); add(b, R, 342, 343); add(b, R, 5, 5)
    # wrap-around spreads from every edge with a tall stack, full-ish board
    rng = np.random.default_rng(7)
    for _ in range(6):
        b = rng.integers(-6, 7, size=49)
        add(b, R, 100, 343); add(b, B, 101, 343)
    # all-own board (294 spread actions, no spawn)
    add(np.full(49, 1), R, 60, 343)
    add(np.full(49, -1), B, 61, 343)


def make_rules_golden():
    rng = np.random.default_rng(2023)
    rows = []
    play_random_games(rng, 3, 343, rows)
    play_random_games(rng, 6, 40, rows)
    play_random_games(rng, 10, 7, rows)
    crafted_positions(rows)
    out = {k: np.array([r[k] for r in rows]) for k in rows[0]}
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(HERE, "rules_transitions.npz"), **out)
    print("rules transitions:", len(rows), "outcomes:", np.bincount(out["next_outcome"], minlength=4))


def make_symmetry_golden():
    g = InflexionGame(7, max_turns=343)
    planes = np.arange(4 * 49).reshape(4, 7, 7)
    policy = np.arange(7 * 49).reshape(7, 7, 7)
    out = dict(
        planes36=np.array(g.symmetries(planes)).astype(np.int16),
        policy36=np.array(g.symmetries(policy)).astype(np.int16),
        rot=np.array([g.rotate(planes, k) for k in range(6)]).astype(np.int16),
        trans=np.array([[g.translate(planes, t, ax) for t in range(7)] for ax in "rqs"]).astype(np.int16),
        rand=np.array([[[g.translate(g.rotate(planes, k), t, ax) for ax in "rqs"] for t in range(7)]
                       for k in range(6)]).astype(np.int16),
    )
    np.savez_compressed(os.path.join(HERE, "symmetries.npz"), **out)
    print("symmetries:", out["planes36"].shape, out["policy36"].shape, out["rand"].shape)


# ------------------------------------------------------------------------------------------------
# MCTS / self-play with the stub net and injected draws
# ------------------------------------------------------------------------------------------------

class StubNet(NNetWrapper):                       # MCTS.py:22 insists on the subclass
    mode = "float"

    def __init__(self, game):                     # no CNN is built
        self.calls = 0

    def predict(self, board):
        self.calls += 1
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(board))[None], self.mode)
        return pi.numpy()[0], v.numpy()           # f32[343], f32 ndarray of shape (1,)  (NNet.py:94)


class Ctx:
    def __init__(self, seed):
        self.draws = PhiloxDraws(seed)
        self.move, self.sim, self.depth = -1, -1, 0
        self.records = []


def traced_mcts(ctx):
    class TracedMCTS(MCTS):
        def getActionProb(self, game, temp=1):
            ctx.move += 1
            ctx.sim = -1
            probs = super().getActionProb(game, temp)
            s = game.to_planes().tobytes()
            counts = np.array([self.Nsa.get((s, a), 0) for a in range(343)], dtype=np.int32)
            q = np.array([float(np.asarray(self.Qsa[(s, a)]).reshape(-1)[0]) if (s, a) in self.Qsa else 0.0
                          for a in range(343)], dtype=np.float64)
            q_is32 = np.array([isinstance(self.Qsa.get((s, a)), (np.ndarray, np.floating))
                               for a in range(343)], dtype=np.uint8)
            ctx.records.append(dict(counts=counts, q=q, q_is32=q_is32, probs=np.asarray(probs, dtype=np.float64),
                                    ns=self.Ns[s], n_nodes=len(self.Ps), temp=temp,
                                    board=game._board.reshape(49).astype(np.int8), player=game.player.num,
                                    turn=game._curr_turn))
            return probs

        def search(self, game):
            if ctx.depth == 0:
                ctx.sim += 1
            ctx.depth += 1
            try:
                return super().search(game)
            finally:
                ctx.depth -= 1

        def reset(self):
            return TracedMCTS(self.nnet, self.args)
    return TracedMCTS


class PatchedRandom:
    """Replace the numpy global-RNG call sites on the path with injected draws."""

    def __init__(self, ctx):
        self.ctx = ctx

    def __enter__(self):
        self.saved = (np.random.randint, np.random.choice)
        ctx = self.ctx

        def randint(lo, hi=None, size=None):
            assert size is None and lo == 0 and hi in (6, 7)
            k, t, _ = ctx.draws.symmetry(ctx.move, ctx.sim)
            return int(k if hi == 6 else t)

        def choice(a, size=None, replace=True, p=None):
            assert size is None
            if isinstance(a, list):                                  # InflexionGame.py:121
                return a[ctx.draws.symmetry(ctx.move, ctx.sim)[2]]
            if p is not None:                                        # Coach.py:81
                act = sample_action(np.asarray(p), ctx.draws.uniform(ctx.move))
                ctx.records[-1]["action"] = act
                return act
            a = np.asarray(a)                                        # MCTS.py:53
            return a[ctx.draws.tiebreak(ctx.move, len(a))]

        np.random.randint, np.random.choice = randint, choice
        return self

    def __exit__(self, *exc):
        np.random.randint, np.random.choice = self.saved


def check_choice_equivalence():
    """np.random.choice(n, p=pi) == searchsorted(cumsum(pi)/cumsum[-1], u, 'right') with u = the next
    random_sample() of the same stream."""
    rs = np.random.RandomState(5)
    for trial in range(2000):
        n_pos = rs.randint(1, 40)
        c = np.zeros(343)
        c[rs.choice(343, n_pos, replace=False)] = rs.randint(1, 50, n_pos)
        pi = c / c.sum()
        if trial % 3 == 0:
            pi = np.zeros(343, dtype=np.int8); pi[rs.randint(343)] = 1
        state = rs.get_state()
        got = rs.choice(343, p=pi)
        rs2 = np.random.RandomState(); rs2.set_state(state)
        assert got == sample_action(pi, rs2.random_sample())
    print("np.random.choice restatement: 2000/2000 equal")


def run_episode(name, seed, sims, cpuct, temp_threshold, max_turns, mode, max_moves=None):
    game = InflexionGame(7, max_turns=max_turns, max_power=6)
    StubNet.mode = mode
    nnet = StubNet(game)
    args = dotdict({"numMCTSSims": sims, "cpuct": cpuct, "tempThreshold": temp_threshold})
    ctx = Ctx(seed)
    Traced = traced_mcts(ctx)
    coach = Coach(game, nnet, args)
    examples = None
    with PatchedRandom(ctx):
        mcts = Traced(nnet, args)
        if max_moves is None:
            examples = coach.executeEpisode((game.restarted(), mcts))
        else:                                   # the same loop as Coach.py:66-82, cut after max_moves
            g = game.restarted()
            for step in range(1, max_moves + 1):
                pi = mcts.getActionProb(g, temp=int(step < temp_threshold))
                g = g.to_next_state(int(np.random.choice(len(pi), p=pi)))
                if g.outcome != GameOutcome.ONGOING:
                    break
    rec = ctx.records
    out = dict(
        seed=seed, sims=sims, cpuct=float(cpuct), cpuct_is_int=isinstance(cpuct, int),
        temp_threshold=temp_threshold, max_turns=max_turns, mode=np.array(mode),
        counts=np.array([r["counts"] for r in rec]), q=np.array([r["q"] for r in rec]),
        q_is32=np.array([r["q_is32"] for r in rec]), probs=np.array([r["probs"] for r in rec]),
        ns=np.array([r["ns"] for r in rec]), n_nodes=np.array([r["n_nodes"] for r in rec]),
        temp=np.array([r["temp"] for r in rec]), action=np.array([r["action"] for r in rec]),
        board=np.array([r["board"] for r in rec]), player=np.array([r["player"] for r in rec]),
        turn=np.array([r["turn"] for r in rec]), net_calls=nnet.calls,
        numpy_version=np.array(np.__version__))
    if examples is not None:
        out["ex_planes"] = np.array([e[0] for e in examples]).astype(np.int16)
        out["ex_pi"] = np.array([e[1] for e in examples], dtype=np.float64)
        out["ex_v"] = np.array([e[2] for e in examples], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, f"mcts_{name}.npz"), **out)
    f64_edges = int((1 - out["q_is32"])[out["counts"] > 0].sum())
    print(f"mcts_{name}: moves={len(rec)} net_calls={nnet.calls} nodes={rec[-1]['n_nodes']} "
          f"py-f64 root edges={f64_edges} examples={0 if examples is None else len(examples)}")


if __name__ == "__main__":
    print("numpy", np.__version__, "torch", torch.__version__)
    check_choice_equivalence()
    make_rules_golden()
    make_symmetry_golden()
    run_episode("a_float_25", seed=11, sims=25, cpuct=1, temp_threshold=30, max_turns=343, mode="float", max_moves=48)
    run_episode("b_dyadic_50", seed=12, sims=50, cpuct=1.25, temp_threshold=6, max_turns=343, mode="dyadic", max_moves=20)
    run_episode("c_short_full", seed=13, sims=40, cpuct=1.0, temp_threshold=4, max_turns=9, mode="float")
    run_episode("d_short_full", seed=14, sims=100, cpuct=2, temp_threshold=30, max_turns=5, mode="float")
