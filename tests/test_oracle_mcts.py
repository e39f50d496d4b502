"""The MCTS oracle against root visit counts / Q bit patterns / actions frozen from the live reference's
MCTS + Coach (tests/golden/make_golden.py), using the deterministic stub net and injected Philox draws."""
import os

import numpy as np
import pytest
import torch

from inflexion_b200.stubnet import stub_policy_value
from oracle.inflexion_mcts import (OracleMCTS, PhiloxDraws, pairwise_sum_f32, philox4x32, sample_action,
                                   self_play_moves)

EPISODES = ["a_float_25", "b_dyadic_50", "c_short_full", "d_short_full"]


def stub_predict(mode):
    def predict(planes):
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None], mode)
        return pi.numpy()[0], v.numpy()[0]
    return predict


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 127, 128, 129, 168, 175, 343, 343, 343):
        a = (rng.random(n) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)
        assert pairwise_sum_f32(a).tobytes() == a.sum().tobytes()
    for _ in range(200):
        a = (rng.random(343) < 0.25) * rng.random(343)
        a = a.astype(np.float32)
        assert pairwise_sum_f32(a).tobytes() == a.sum().tobytes()


def test_sample_action_edges():
    pi = np.zeros(343); pi[5] = 0.5; pi[300] = 0.5
    assert sample_action(pi, 0.0) == 5 and sample_action(pi, 0.4999) == 5
    assert sample_action(pi, 0.5) == 300 and sample_action(pi, 0.999999) == 300
    onehot = np.zeros(343, dtype=np.int8); onehot[17] = 1
    assert sample_action(onehot, 0.0) == 17 and sample_action(onehot, 0.99) == 17


@pytest.mark.parametrize("name", EPISODES)
def test_oracle_reproduces_reference_episode(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    cpuct = int(g["cpuct"]) if bool(g["cpuct_is_int"]) else float(g["cpuct"])
    draws = PhiloxDraws(int(g["seed"]))
    n_moves = len(g["action"])
    it = self_play_moves(stub_predict(str(g["mode"])), int(g["sims"]), cpuct, draws,
                         temp_threshold=int(g["temp_threshold"]), max_moves=n_moves, max_turns=int(g["max_turns"]))
    last = None
    for i, rec in enumerate(it):
        assert np.array_equal(rec["counts"], g["counts"][i]), f"visit counts differ at move {i}"
        assert np.array_equal(np.asarray(rec["pi"], dtype=np.float64), g["probs"][i])
        assert rec["action"] == g["action"][i]
        mcts, s = rec["mcts"], None
        # Q bit patterns and kinds of the root edges
        root = rec["planes"].tobytes()
        for a in np.nonzero(g["counts"][i])[0]:
            q = mcts.Q[(root, int(a))]
            assert q.is32 == bool(g["q_is32"][i][a])
            assert np.float64(q.x).tobytes() == g["q"][i][a].tobytes()
        assert mcts.Ns[root] == g["ns"][i] and len(mcts.P) == g["n_nodes"][i]
        last = rec
    assert i == n_moves - 1
    if "ex_v" in g.files:          # the episode ran to the end
        assert last["game"].outcome != 0


def test_ndarray_sum_and_handwritten_sum_agree_in_search(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts_a_float_25.npz"))
    draws = PhiloxDraws(int(g["seed"]))
    for i, rec in enumerate(self_play_moves(stub_predict("float"), 25, 1, draws, max_moves=6, exact_sum=False)):
        assert np.array_equal(rec["counts"], g["counts"][i])


def test_float32_sqrt_identity_behind_the_node_header():
    """csrc/ifx_tree.cuh:ifx_head_of computes both PUCT square roots (MCTS.py:122,125) as ONE float32 square root.
    Valid iff, for every visit count the 17-bit N field can hold, float32(math.sqrt(Ns)) and
    float32(math.sqrt(Ns + 1e-8)) equal the correctly rounded float32 square root of Ns."""
    n = np.arange(0, 1 << 17, dtype=np.int64)
    f32 = np.sqrt(n.astype(np.float32))
    assert np.array_equal(np.sqrt(n.astype(np.float64)).astype(np.float32), f32)
    assert np.array_equal(np.sqrt(n.astype(np.float64) + 1e-8).astype(np.float32)[1:], f32[1:])
    assert np.float32(np.sqrt(np.float64(0) + 1e-8)) == np.float32(1e-4)
