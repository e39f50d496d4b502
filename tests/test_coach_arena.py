"""Callers of the hot path with the reference's interfaces: Coach.executeEpisode, Arena, example files."""
import os
import pickle

import numpy as np
import pytest
import torch

import inflexion_b200.backend as backend_mod
from inflexion_b200.Arena import Arena
from inflexion_b200.Coach import Coach
from inflexion_b200.flags import PlayerColour
from inflexion_b200.inflexion.InflexionGame import InflexionGame
from inflexion_b200.inflexion.InflexionPlayers import GreedyPlayer, MCTSPlayer, RandomPlayer
from inflexion_b200.MCTS import MCTS
from inflexion_b200.stubnet import stub_policy_value
from inflexion_b200.utils import dotdict
from oracle.inflexion_mcts import PhiloxDraws, sample_action


@pytest.fixture(autouse=True)
def use_backend(backend, monkeypatch):
    monkeypatch.setattr(backend_mod, "_DEFAULT", backend)
    return backend


class StubNet:
    def __init__(self, mode="float"):
        self.mode = mode

    def predict(self, board):
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(board))[None], self.mode)
        return pi.numpy()[0], v.numpy()

    def predict_batch(self, planes):
        return stub_policy_value(planes, self.mode)


@pytest.mark.parametrize("name", ["c_short_full", "d_short_full"])
def test_execute_episode_reproduces_reference_examples(golden_dir, name, monkeypatch):
    """Coach.executeEpisode over the drop-in classes == the examples the reference's Coach produced (same stub net,
    same injected draws: the action sample of Coach.py:81 is replayed from the Philox uniform the golden run used)."""
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    cpuct = int(g["cpuct"]) if bool(g["cpuct_is_int"]) else float(g["cpuct"])
    args = dotdict({"numMCTSSims": int(g["sims"]), "cpuct": cpuct, "tempThreshold": int(g["temp_threshold"])})
    game = InflexionGame(7, max_turns=int(g["max_turns"]), max_power=6)
    draws, move = PhiloxDraws(int(g["seed"])), [0]

    def choice(a, size=None, replace=True, p=None):
        act = sample_action(np.asarray(p), draws.uniform(move[0]))
        move[0] += 1
        return act
    monkeypatch.setattr(np.random, "choice", choice)
    coach = Coach(game, StubNet(str(g["mode"])), args)
    ex = coach.executeEpisode((game.restarted(), MCTS(StubNet(str(g["mode"])), args, seed=int(g["seed"]))))
    assert len(ex) == len(g["ex_v"])
    assert np.array_equal(np.array([e[0] for e in ex]), g["ex_planes"].astype(np.int64))
    assert np.array([e[1] for e in ex], dtype=np.float64).tobytes() == g["ex_pi"].tobytes()
    assert np.array([e[2] for e in ex], dtype=np.float64).tobytes() == g["ex_v"].tobytes()
    assert isinstance(ex[0][1], list) and ex[0][0].dtype == np.int64 and ex[0][0].shape == (4, 7, 7)


def test_arena_results_and_colour_split():
    game = InflexionGame(7, max_turns=12, max_power=6)
    np.random.seed(1)
    arena = Arena(GreedyPlayer(), RandomPlayer(), game)
    wins, draws = arena.playGame(PlayerColour.RED, PlayerColour.BLUE)
    assert sum(wins.values()) + draws == 1 and all(isinstance(k, PlayerColour) for k in wins)
    one, two, dr = arena.playGames(4)
    assert one + two + dr == 4
    mp = MCTSPlayer(MCTS(StubNet(), dotdict({"numMCTSSims": 6, "cpuct": 1.0}), seed=2))
    one, two, dr = Arena(mp, RandomPlayer(), game).playGames(2)
    assert one + two + dr == 2


def test_example_file_format_roundtrip(tmp_path):
    game = InflexionGame(7, max_turns=4, max_power=6)
    args = dotdict({"numMCTSSims": 5, "cpuct": 1, "tempThreshold": 2, "checkpoint": str(tmp_path),
                    "load_folder_file": (str(tmp_path), "checkpoint_0.pth.tar")})
    coach = Coach(game, StubNet(), args)
    np.random.seed(0)
    ex = coach.executeEpisode((game.restarted(), MCTS(StubNet(), args, seed=1)))
    from collections import deque
    coach.trainExamplesHistory.append(deque(ex, maxlen=1000))
    coach.saveTrainExamples(0)
    path = os.path.join(tmp_path, "checkpoint_0.pth.tar.examples")
    with open(path, "rb") as f:
        hist = pickle.load(f)                                      # plain pickle, as the reference writes it
    assert isinstance(hist, list) and isinstance(hist[0], deque) and len(hist[0]) == len(ex)
    board, pi, v = hist[0][0]
    assert board.shape == (4, 7, 7) and len(pi) == 343 and isinstance(v, float)
    c2 = Coach(game, StubNet(), args)
    c2.loadTrainExamples()
    assert c2.skipFirstSelfPlay and len(c2.trainExamplesHistory[0]) == len(ex)


def test_greedy_kernel_equals_greedy_player(backend):
    from inflexion_b200 import rules
    from inflexion_b200.state import pack_states
    from oracle import inflexion_rules as R
    b, p, t, _ = R.random_reachable_states(60, seed=8)
    st = pack_states(b, p, t)
    got = rules.greedy_actions(backend.states_tensor(st), backend=backend).cpu().numpy()
    for i in range(60):
        g = InflexionGame(7, curr_player=PlayerColour.RED if p[i] > 0 else PlayerColour.BLUE,
                          board=b[i].astype(int).reshape(7, 7), curr_turn=int(t[i]), max_turns=343)
        assert got[i] == GreedyPlayer().play(g)


def test_batched_arena(backend):
    from inflexion_b200.arena_batched import BatchedArena
    ev = lambda x: stub_policy_value(x)  # noqa: E731
    for opp in ("random", "greedy"):
        arena = BatchedArena(ev, num_sims=6, opponent=opp, seed=3, max_turns=10, backend=backend)
        one, two, dr = arena.playGames(7)
        assert one + two + dr == 7
        assert (one, two, dr) == BatchedArena(ev, num_sims=6, opponent=opp, seed=3, max_turns=10, backend=backend).playGames(7)


# ---- Arena against the reference-derived fixtures (tests/golden/make_golden_arena.py) ------------------------------------
@pytest.mark.parametrize("name", ["vs_greedy", "vs_random"])
def test_batched_arena_reproduces_the_reference_arena(backend, golden_dir, name):
    """The live reference's Arena.playGame (Arena.py:38-88) with its MCTSPlayer on the stub net against its GreedyPlayer /
    RandomPlayer, draws injected, frozen move by move: BatchedArena plays the same games -- every move of both players,
    every result, the W/L/D totals, with the reference's colour split (i <= num // 2)."""
    from inflexion_b200.arena_batched import BatchedArena
    g = np.load(os.path.join(golden_dir, f"arena_{name}.npz"))
    mode = str(g["mode"])
    arena = BatchedArena(lambda x: stub_policy_value(x, mode), int(g["sims"]), float(g["cpuct"]), opponent=str(g["opponent"]),
                         seed=int(g["seed"]), max_turns=int(g["max_turns"]), backend=backend, record=True)
    totals = arena.playGames(int(g["num"]), shard=False)
    got = arena.last_actions.T                                  # [games, plies]
    want = g["actions"]
    assert got.shape[1] >= want.shape[1]
    assert np.array_equal(got[:, :want.shape[1]], want) and bool((got[:, want.shape[1]:] == -1).all())
    assert np.array_equal(arena.last_results, g["result"])
    assert list(totals) == g["totals"].tolist()


@pytest.mark.parametrize("name", ["vs_greedy", "vs_random"])
def test_object_arena_reproduces_the_reference_arena(backend, golden_dir, name, monkeypatch):
    """The same fixtures through the drop-in classes a pit.py user touches: Arena.playGames over MCTSPlayer(MCTS(...)),
    GreedyPlayer / RandomPlayer objects (one game at a time).  RandomPlayer draws with np.random.choice, as the
    reference's does; the test injects the fixture's draws there."""
    from oracle.inflexion_mcts import philox4x32
    g = np.load(os.path.join(golden_dir, f"arena_{name}.npz"))
    monkeypatch.setattr(backend_mod, "_DEFAULT", backend)
    seed, mode, num = int(g["seed"]), str(g["mode"]), int(g["num"])

    class Net:
        def predict(self, planes):
            pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None], mode)
            return pi.numpy()[0], v.numpy()
    args = dotdict({"numMCTSSims": int(g["sims"]), "cpuct": float(g["cpuct"])})
    game = InflexionGame(7, max_turns=int(g["max_turns"]), max_power=6)
    state = {"game": -1, "ply": 0}
    played = [[] for _ in range(num)]

    class Spy:
        def __init__(self, inner, counts_games):
            self.inner, self.counts_games = inner, counts_games

        def play(self, gm):
            state["ply"] = gm._curr_turn
            a = self.inner.play(gm)
            played[state["game"]].append(int(a))
            return a

        def reset(self):
            if self.counts_games:
                state["game"] += 1
            return Spy(self.inner.reset(), False)
    key = ((seed ^ 0x5EED) & 0xFFFFFFFF, ((seed ^ 0x5EED) >> 32) & 0xFFFFFFFF)

    def choice(a, *rest, **kw):
        a = np.asarray(a)
        return a[philox4x32((state["game"], state["ply"], 0, 3), key)[0] % len(a)]
    monkeypatch.setattr(np.random, "choice", choice)
    opp = GreedyPlayer() if str(g["opponent"]) == "greedy" else RandomPlayer()
    arena = Arena(Spy(MCTSPlayer(MCTS(Net(), args, seed=seed, backend=backend)), True), Spy(opp, False), game)
    totals = arena.playGames(num)
    assert list(totals) == g["totals"].tolist()
    for i in range(num):
        assert played[i] == g["actions"][i, :g["plies"][i]].tolist(), i
