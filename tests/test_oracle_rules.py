"""The rules oracle against the golden transitions frozen from the live reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import inflexion_rules as R


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rules_transitions.npz"))


def test_golden_covers_the_corners(gold):
    assert len(gold["action"]) > 4000
    assert set(np.unique(gold["next_outcome"])) == {0, 1, 2, 3}
    assert (gold["action"] >= 294).any() and (gold["action"] < 294).any()
    assert (np.abs(gold["board"].astype(int)).sum(axis=1) > 48).any()      # spawn forbidden somewhere


def test_scalar_oracle_matches_reference_transitions(gold):
    for i in range(len(gold["action"])):
        g = R.OracleGame(gold["board"][i].astype(np.int64).reshape(7, 7).copy(), int(gold["player"][i]),
                         int(gold["turn"][i]), R.ONGOING, int(gold["max_turns"][i]))
        mask = g.valid_actions_mask()
        assert np.array_equal(np.packbits(mask.astype(np.uint8)), gold["mask"][i])
        assert np.array_equal(g.to_planes().reshape(4, 49), gold["planes"][i])
        assert g.score() == gold["score"][i]
        n = g.to_next_state(int(gold["action"][i]))
        assert np.array_equal(n.board.reshape(49), gold["next_board"][i])
        assert (n.outcome, n.player, n.turn) == (gold["next_outcome"][i], gold["next_player"][i], gold["next_turn"][i])
        assert np.array_equal(g.board.reshape(49), gold["board"][i])       # receiver untouched


def test_batched_oracle_matches_reference_transitions(gold):
    b, p, t = gold["board"], gold["player"].astype(np.int64), gold["turn"].astype(np.int64)
    mask = R.batch_valid_mask(b, p)
    assert np.array_equal(np.packbits(mask, axis=1), gold["mask"])
    assert np.array_equal(R.batch_planes(b, p, t).reshape(-1, 4, 49), gold["planes"])
    for mt in np.unique(gold["max_turns"]):
        sel = gold["max_turns"] == mt
        nb, npl, nt, out = R.batch_step(b[sel], p[sel], t[sel], gold["action"][sel], int(mt))
        assert np.array_equal(nb, gold["next_board"][sel])
        assert np.array_equal(npl, gold["next_player"][sel])
        assert np.array_equal(nt, gold["next_turn"][sel])
        assert np.array_equal(out, gold["next_outcome"][sel])


def test_symmetries_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "symmetries.npz"))
    planes = np.arange(4 * 49).reshape(4, 7, 7)
    policy = np.arange(7 * 49).reshape(7, 7, 7)
    assert np.array_equal(np.array(R.symmetries(planes)), g["planes36"])
    assert np.array_equal(np.array(R.symmetries(policy)), g["policy36"])
    for k in range(6):
        assert np.array_equal(R.rotate(planes, k), g["rot"][k])
        for t in range(7):
            for ax in range(3):
                assert np.array_equal(R.random_symmetry(planes, k, t, ax), g["rand"][k, t, ax])
                assert np.array_equal(R.translate(planes, t, "rqs"[ax]), g["trans"][ax, t])
    assert len({x.tobytes() for x in R.symmetries(planes)}) == 36


def test_random_reachable_states_are_consistent():
    b, p, t, a = R.random_reachable_states(300, seed=3)
    assert b.shape == (300, 49) and np.abs(b).max() <= 6
    mask = R.batch_valid_mask(b, p)
    assert mask[np.arange(300), a].all()
    # red moves on even turns
    assert np.array_equal(p, np.where(t % 2 == 0, 1, -1))


@pytest.mark.skipif(not os.path.isdir("/root/reference/inflexion"), reason="live reference not mounted")
def test_oracle_against_live_reference_random_play():
    import sys
    sys.path.insert(0, "/root/reference")
    from inflexion.InflexionGame import InflexionGame
    from flags import GameOutcome
    code = {GameOutcome.ONGOING: 0, GameOutcome.DRAW: 1, GameOutcome.WON: 2, GameOutcome.LOST: 3}
    rng = np.random.default_rng(99)
    ref = InflexionGame(7, max_turns=60)
    mine = R.OracleGame(max_turns=60)
    while ref.outcome == GameOutcome.ONGOING:
        m = ref.valid_actions_mask()
        assert np.array_equal(m, mine.valid_actions_mask())
        assert ref.to_planes().tobytes() == mine.key()
        a = int(rng.choice(np.nonzero(m)[0]))
        ref, mine = ref.to_next_state(a), mine.to_next_state(a)
        assert np.array_equal(ref._board, mine.board) and code[ref.outcome] == mine.outcome
