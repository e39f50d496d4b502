"""Golden fixtures for the Arena path from the LIVE reference:

    python tests/golden/make_golden_arena.py

Runs the reference's own ``Arena.playGame`` (Arena.py:38-88) with its own ``MCTSPlayer`` / ``GreedyPlayer`` /
``RandomPlayer`` (inflexion/InflexionPlayers.py:24-34,61-88), unmodified, for the games ``Arena.playGames`` would play
(colour split ``i <= num // 2``, Arena.py:126-129 -- its process pool is not used: patched draws do not cross into
worker processes), the MCTS player on the deterministic stub net, and numpy's global-RNG call sites replaced by the
Philox draws the device uses:
  * leaf symmetries and the arg-max tie-break of game i's MCTS player: keyed (seed; game i, its k-th move, simulation);
  * RandomPlayer's ``np.random.choice(actions)``: index = Philox(seed ^ 0x5EED; game i, ply, 0, 3).w0 % len(actions)
    (the draw of ifx_rules_random_actions).
Freezes every game's move sequence and result and the W/L/D totals.  The players are wrapped by an observer that only
records the ply and the action; nothing of the reference is altered.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up sys.path for the reference, StubNet, traced MCTS, PhiloxDraws)

from Arena import Arena  # noqa: E402  (reference)
from flags import PlayerColour  # noqa: E402  (reference)
from inflexion.InflexionGame import InflexionGame  # noqa: E402  (reference)
from inflexion.InflexionPlayers import GreedyPlayer, MCTSPlayer, Player, RandomPlayer  # noqa: E402  (reference)
from utils import dotdict  # noqa: E402  (reference)
from oracle.inflexion_mcts import philox4x32  # noqa: E402

_MASK = 0xFFFFFFFF


class Spy(Player):
    """Observes which ply is being played and what was played; delegates everything to the reference player."""

    def __init__(self, inner, ctx, is_mcts):
        super().__init__()
        self.inner, self.ctx, self.is_mcts = inner, ctx, is_mcts

    def play(self, game):
        self.ctx.ply, self.ctx.in_mcts = game._curr_turn, self.is_mcts
        a = self.inner.play(game)
        self.ctx.in_mcts = False
        self.ctx.actions.append(int(a))
        return a

    def reset(self):
        return Spy(self.inner.reset(), self.ctx, self.is_mcts)


class ArenaRandom(mg.PatchedRandom):
    def __init__(self, ctx, seed, game_index):
        super().__init__(ctx)
        self.seed, self.game_index = seed, game_index

    def __enter__(self):
        super().__enter__()
        inner_choice, ctx = np.random.choice, self.ctx
        key = ((self.seed ^ 0x5EED) & _MASK, ((self.seed ^ 0x5EED) >> 32) & _MASK)

        def choice(a, size=None, replace=True, p=None):
            if p is None and not isinstance(a, list) and not ctx.in_mcts:      # RandomPlayer (InflexionPlayers.py:30)
                a = np.asarray(a)
                w = philox4x32((self.game_index & _MASK, ctx.ply & _MASK, 0, 3), key)
                return a[w[0] % len(a)]
            return inner_choice(a, size, replace, p)
        np.random.choice = choice
        return self


def run_arena(name, opponent, num, sims, cpuct, max_turns, seed, mode="float"):
    game = InflexionGame(7, max_turns=max_turns, max_power=6)
    mg.StubNet.mode = mode
    nnet = mg.StubNet(game)
    args = dotdict({"numMCTSSims": sims, "cpuct": cpuct})
    red, blue = list(PlayerColour)
    actions, results, first_red, plies = [], [], [], []
    totals = {red: 0, blue: 0, "draws": 0}
    for i in range(num):
        ctx = mg.Ctx(seed)
        ctx.draws = mg.PhiloxDraws(seed, game=i)
        ctx.ply, ctx.in_mcts, ctx.actions = 0, False, []
        Traced = mg.traced_mcts(ctx)
        opp = GreedyPlayer() if opponent == "greedy" else RandomPlayer()
        arena = Arena(Spy(MCTSPlayer(Traced(nnet, args)), ctx, True), Spy(opp, ctx, False), game)
        first, second = (red, blue) if i <= num // 2 else (blue, red)           # Arena.py:126-129
        with ArenaRandom(ctx, seed, i):
            wins, draws = arena.playGame(first, second)
        totals[red] += wins[red]; totals[blue] += wins[blue]; totals["draws"] += draws
        results.append(1 if wins[red] else (-1 if wins[blue] else 0))
        first_red.append(first is red)
        plies.append(len(ctx.actions))
        actions.append(ctx.actions)
    width = max(plies)
    act = np.full((num, width), -1, dtype=np.int32)
    for i, a in enumerate(actions):
        act[i, :len(a)] = a
    out = dict(opponent=np.array(opponent), num=num, sims=sims, cpuct=float(cpuct), max_turns=max_turns, seed=seed,
               mode=np.array(mode), actions=act, plies=np.array(plies), first_red=np.array(first_red),
               result=np.array(results), totals=np.array([totals[red], totals[blue], totals["draws"]]),
               numpy_version=np.array(np.__version__))
    np.savez_compressed(os.path.join(HERE, f"arena_{name}.npz"), **out)
    print(f"arena_{name}: {num} games, plies {plies}, results {results}, totals (MCTS, {opponent}, draws) = {out['totals'].tolist()}")


if __name__ == "__main__":
    run_arena("vs_greedy", "greedy", num=7, sims=16, cpuct=1.0, max_turns=16, seed=31)
    run_arena("vs_random", "random", num=6, sims=12, cpuct=1.0, max_turns=22, seed=32)
