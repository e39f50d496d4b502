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


def test_continuous_selfplay_restarts_finished_games(backend):
    """auto_restart: a finished game starts its next episode at once on a fresh RNG stream; every finished episode
    yields its examples, and the first episode of each slot is the same game a run without restarts plays."""
    kw = dict(cpuct=1.0, temp_threshold=3, seed=4, max_turns=4, game_id_base=40, backend=backend)
    base = SelfPlayEngine(4, 10, lambda x: stub_policy_value(x), **kw)
    n0 = base.play_episode()
    eng = SelfPlayEngine(4, 10, lambda x: stub_policy_value(x), auto_restart=True, restart_stride=1000, record_moves=16, **kw)
    for _ in range(12):
        eng.play_move()
    eng.mcts.raise_on_faults()
    assert np.array_equal(eng.ex_actions[:n0].cpu().numpy()[base.ex_actions[:n0].cpu().numpy() >= 0],
                          base.ex_actions[:n0].cpu().numpy()[base.ex_actions[:n0].cpu().numpy() >= 0])
    done = int(eng.restarts[0])
    assert done >= 8                                            # games last at most 5 moves here
    tags = eng.mcts.roots().cpu().numpy().view(np.uint32).reshape(4, 16)[:, 12]
    assert all(int(t) >= 40 + 2000 for t in tags)               # every slot is on its third stream or later
    planes, pi, v = eng.examples(expand_symmetries=False)
    ep_first = (planes[:, 2, 0, 0] == 0).sum()                  # one turn-0 example per finished episode
    assert ep_first == done and len(planes) >= done * 2 and np.allclose(pi.sum(axis=1), 1.0)
    p0, pi0, v0 = base.examples(expand_symmetries=False)
    assert len(p0) > 0
