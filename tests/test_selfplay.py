"""SelfPlayEngine (batched Coach.executeEpisode) against the examples the reference's Coach produced."""
import os

import numpy as np
import pytest

from inflexion_b200.selfplay import SelfPlayEngine
from inflexion_b200.stubnet import stub_policy_value


@pytest.mark.parametrize("name", ["c_short_full", "d_short_full"])
def test_episode_examples_match_reference_coach(backend, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    cpuct = int(g["cpuct"]) if bool(g["cpuct_is_int"]) else float(g["cpuct"])
    eng = SelfPlayEngine(1, int(g["sims"]), lambda x: stub_policy_value(x, str(g["mode"])), cpuct=cpuct,
                         temp_threshold=int(g["temp_threshold"]), seed=int(g["seed"]), max_turns=int(g["max_turns"]),
                         backend=backend)
    moves = eng.play_episode()
    assert moves == len(g["action"])
    assert np.array_equal(eng.ex_actions[:moves, 0].cpu().numpy(), g["action"])
    planes, pi, v = eng.examples()
    assert np.array_equal(planes, g["ex_planes"].astype(np.int64))
    assert pi.tobytes() == g["ex_pi"].tobytes()
    assert v.tobytes() == g["ex_v"].tobytes()
    # the aligned variant differs from the reference's but has the same magnitude
    _, _, v_fixed = eng.examples(fix_alignment=True)
    assert np.array_equal(np.abs(v_fixed), np.abs(v)) and not np.array_equal(v_fixed, v)


def test_many_games_lockstep_and_restart(backend):
    eng = SelfPlayEngine(5, 12, lambda x: stub_policy_value(x), cpuct=1.0, temp_threshold=3, seed=4, max_turns=6,
                         game_id_base=40, backend=backend)
    moves = eng.play_episode()
    st = eng.mcts.status().cpu().numpy()
    assert (st[:, 4] != 0).all() and (st[:, 2] == 0).all() and moves <= 7
    first = eng.ex_actions[:moves].cpu().numpy().copy()
    planes, pi, v = eng.examples(expand_symmetries=False)
    assert len(planes) == int((first >= 0).sum()) and np.allclose(pi.sum(axis=1), 1.0)
    eng.restart()
    assert eng.play_episode() == moves
    assert np.array_equal(eng.ex_actions[:moves].cpu().numpy(), first)        # same seed, same game ids -> same games
