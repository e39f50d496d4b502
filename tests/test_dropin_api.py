"""The reference-shaped object API (InflexionGame / MCTS / players) -- same names, argument meaning, return
dtypes and error behaviour as the reference's classes, computed by the kernels."""
import os
import pickle

import numpy as np
import pytest
import torch

import inflexion_b200.backend as backend_mod
from inflexion_b200.flags import GameOutcome, PlayerColour
from inflexion_b200.inflexion.InflexionGame import InflexionGame, Move
from inflexion_b200.inflexion.InflexionPlayers import GreedyPlayer, MCTSPlayer, RandomPlayer
from inflexion_b200.MCTS import MCTS
from inflexion_b200.stubnet import stub_policy_value
from inflexion_b200.utils import dotdict
from oracle import inflexion_rules as R

CODE = {GameOutcome.ONGOING: 0, GameOutcome.DRAW: 1, GameOutcome.WON: 2, GameOutcome.LOST: 3}


@pytest.fixture(autouse=True)
def use_backend(backend, monkeypatch):
    monkeypatch.setattr(backend_mod, "_DEFAULT", backend)
    return backend


class StubNet:
    def __init__(self, mode="float"):
        self.mode, self.calls = mode, 0

    def predict(self, board):
        assert board.dtype == np.int64 and board.shape == (4, 7, 7)
        self.calls += 1
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(board))[None], self.mode)
        return pi.numpy()[0], v.numpy()


def test_game_object_semantics():
    g = InflexionGame(7, max_turns=343, max_power=6)
    assert g.board_shape == (7, 7) and g.policy_shape == (7, 7, 7) and g.max_actions == 343
    assert g.player is PlayerColour.RED and g.outcome is GameOutcome.ONGOING and g._curr_turn == 0
    m = g.valid_actions_mask()
    assert m.dtype == np.int64 and m.shape == (343,) and m.sum() == 49 and m[294:].all()
    pl = g.to_planes()
    assert pl.dtype == np.int64 and pl.shape == (4, 7, 7) and len(pl.tobytes()) == 1568
    g2 = g.to_next_state(g.move_to_action((Move.SPAWN, 3, 4)))
    assert g2 is not g and g._curr_turn == 0 and not g._board.any()          # receiver untouched
    assert g2._board[3, 4] == 1 and g2.player is PlayerColour.BLUE and g2._curr_turn == 1
    assert g2.action_to_move(6 * 49 + 3 * 7 + 4) == (Move.SPAWN, 3, 4)
    with pytest.raises(ValueError):
        g2.to_next_state(6 * 49 + 3 * 7 + 4)                                  # occupied cell
    with pytest.raises(ValueError):
        g.to_next_state(0)                                                     # no stack to spread
    # player setter flips the outcome (Game.py:49-62)
    b = np.zeros((7, 7), dtype=int); b[0, 0] = 6; b[1, 0] = -1
    w = InflexionGame(7, board=b, curr_turn=10, max_turns=343).to_next_state(0)   # RED spreads over BLUE's only stack
    assert w.outcome is GameOutcome.LOST and w.player is PlayerColour.BLUE
    w.player = PlayerColour.RED
    assert w.outcome is GameOutcome.WON
    w.player = PlayerColour.RED
    assert w.outcome is GameOutcome.WON
    # restarted / pickling (Arena ships games to worker processes)
    r = w.restarted()
    assert r._curr_turn == 0 and not r._board.any() and r._max_turns == 343
    c = pickle.loads(pickle.dumps(g2))
    assert np.array_equal(c._board, g2._board) and c.player is g2.player
    with pytest.raises(NotImplementedError):
        InflexionGame(5)


def test_random_playout_against_oracle_objects():
    rng = np.random.default_rng(5)
    g = InflexionGame(7, max_turns=50)
    o = R.OracleGame(max_turns=50)
    while g.outcome is GameOutcome.ONGOING:
        m = g.valid_actions_mask()
        assert np.array_equal(m, o.valid_actions_mask())
        assert g.to_planes().tobytes() == o.key()
        assert g.score() == o.score() and g.total_power() == o.total_power()
        assert g.power_diff(g.player) == o.player * int(o.board.sum())
        a = int(rng.choice(np.flatnonzero(m)))
        g, o = g.to_next_state(a), o.to_next_state(a)
        assert np.array_equal(g._board, o.board) and CODE[g.outcome] == o.outcome
    assert g._curr_turn == 51


def test_symmetry_methods(golden_dir):
    gold = np.load(os.path.join(golden_dir, "symmetries.npz"))
    g = InflexionGame(7)
    planes = np.arange(4 * 49).reshape(4, 7, 7)
    policy = np.arange(7 * 49).reshape(7, 7, 7)
    sy = g.symmetries(planes)
    assert len(sy) == 36 and sy[0].dtype == planes.dtype and np.array_equal(np.array(sy), gold["planes36"])
    assert np.array_equal(np.array(g.symmetries(policy.astype(np.float64))), gold["policy36"].astype(np.float64))
    onehot = np.zeros((7, 7, 7), dtype=np.int8); onehot[6, 2, 5] = 1          # pi at temp 0 is int8 (MCTS.py:54)
    s8 = g.symmetries(onehot)
    assert s8[0].dtype == np.int8 and all(x.sum() == 1 for x in s8)
    for k in range(-2, 8):
        assert np.array_equal(g.rotate(planes, k), gold["rot"][k % 6])
    for ai, ax in enumerate("rqs"):
        for t in range(7):
            assert np.array_equal(g.translate(planes, t, ax), gold["trans"][ai, t])
    np.random.seed(3)
    k, t = np.random.randint(0, 6), np.random.randint(0, 7)
    ax = "rqs".index(np.random.choice(["r", "q", "s"]))
    np.random.seed(3)
    assert np.array_equal(g.random_symmetry(planes), gold["rand"][k, t, ax])   # same draws as the reference's call order


@pytest.mark.parametrize("name", ["c_short_full", "d_short_full"])
def test_mcts_object_reproduces_reference_episode(golden_dir, name):
    """Coach.executeEpisode's loop written against the drop-in classes, checked against the frozen reference run."""
    gd = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    cpuct = int(gd["cpuct"]) if bool(gd["cpuct_is_int"]) else float(gd["cpuct"])
    args = dotdict({"numMCTSSims": int(gd["sims"]), "cpuct": cpuct, "tempThreshold": int(gd["temp_threshold"])})
    net = StubNet(str(gd["mode"]))
    mcts = MCTS(net, args, seed=int(gd["seed"]))
    game = InflexionGame(7, max_turns=int(gd["max_turns"]), max_power=6)
    step = 0
    while True:
        step += 1
        temp = int(step < args.tempThreshold)
        pi = mcts.getActionProb(game, temp=temp)
        assert pi.shape == (343,) and pi.dtype == (np.float64 if temp else np.int8)
        if temp:
            assert pi.tobytes() == gd["probs"][step - 1].tobytes()
        else:   # the tie-break draw is the device's: the one-hot must sit on an arg-max of the reference counts
            assert gd["counts"][step - 1][int(np.argmax(pi))] == gd["counts"][step - 1].max()
            assert np.array_equal(pi, gd["probs"][step - 1].astype(np.int8))
        assert np.array_equal(mcts.dump()["counts"], gd["counts"][step - 1])
        game = game.to_next_state(int(gd["action"][step - 1]))
        if game.outcome is not GameOutcome.ONGOING:
            break
    assert step == len(gd["action"]) and net.calls == int(gd["net_calls"])


def test_players():
    g = InflexionGame(7, max_turns=30)
    np.random.seed(0)
    rp, gp = RandomPlayer(), GreedyPlayer()
    mp = MCTSPlayer(MCTS(StubNet(), dotdict({"numMCTSSims": 8, "cpuct": 1.0}), seed=3))
    assert isinstance(rp.reset(), RandomPlayer) and isinstance(mp.reset(), MCTSPlayer) and mp.reset().mcts is not mp.mcts
    turn = 0
    while g.outcome is GameOutcome.ONGOING:
        player = (rp, gp, mp)[turn % 3]
        a = player.play(g)
        assert isinstance(a, int) and g.valid_actions_mask()[a] == 1
        if player is gp:        # reference rule: best score(), ties -> highest action (InflexionPlayers.py:61-77)
            cands = []
            for b in np.flatnonzero(g.valid_actions_mask()):
                nb = g.to_next_state(int(b))
                nb.player = nb.player.opponent
                cands.append((nb.score(), int(b)))
            cands.sort(reverse=True)
            assert a == cands[0][1]
        g = g.to_next_state(a)
        turn += 1
    assert g._curr_turn <= 31
