"""The leaf-parallel performance mode (virtual loss; csrc/ifx_tree_vl.cuh, ifx_mcts_config.leaves_per_wave).

It is NOT a parity mode -- K leaves per wave change the visit counts by construction -- so it is tested through
(i) K = 1 on the leaf-parallel kernel reproducing the sequential (reference-exact) search bit for bit, (ii) invariants
every PUCT search must keep (Ns = sum of edge counts in every node, the root's visits = simulations, visits only on
valid moves, determinism), (iii) closeness of the root visit distribution to the sequential search's, and (iv) what it
is for: far fewer, larger net calls."""
import numpy as np
import pytest
import torch

from inflexion_b200.mcts_batched import BatchedMCTS
from inflexion_b200.stubnet import stub_policy_value
from test_mcts_kernel import start_states


def _play(backend, leaves, G=6, sims=40, moves=4, seed=21, calls=None):
    m = BatchedMCTS(G, sims, 1.0, seed=seed, backend=backend, leaves_per_wave=leaves)
    m.reset(start_states(G, 343))
    out = []

    def evaluate(x):
        if calls is not None:
            calls.append(x.shape[0])
        return stub_policy_value(x)
    for _ in range(moves):
        m.search(evaluate)
        m.raise_on_faults()
        st = m.root_stats(want_q=True)
        acts = m.sample_actions(0)
        out.append((st["counts"].cpu().numpy().copy(), st["q64"].cpu().numpy().copy(), acts.cpu().numpy().copy()))
        m.advance(acts)
    return m, out


def test_one_leaf_per_wave_reproduces_the_sequential_search(backend):
    _, exact = _play(backend, 1)
    _, vl1 = _play(backend, -1)                                  # the leaf-parallel kernel with K = 1 (test hook)
    for (c0, q0, a0), (c1, q1, a1) in zip(exact, vl1):
        assert np.array_equal(c0, c1) and q0.tobytes() == q1.tobytes() and np.array_equal(a0, a1)


def test_invariants_and_determinism(backend):
    for K in (4, 8):
        m, a = _play(backend, K, moves=3)
        _, b = _play(backend, K, moves=3)
        for (c0, q0, a0), (c1, q1, a1) in zip(a, b):
            assert np.array_equal(c0, c1) and q0.tobytes() == q1.tobytes() and np.array_equal(a0, a1)
        for mv, (counts, _, _) in enumerate(a):
            # the first move expands the root with its first simulation; later roots are reused nodes
            assert (counts.sum(axis=1) >= 39).all() and (counts.sum(axis=1) <= 40 * (mv + 1)).all()
        tv = m.table_views()
        for g in range(tv["key"].shape[0]):
            used = np.nonzero(tv["key"][g][:, 0] != np.uint64(2 ** 64 - 1))[0]
            assert len(used) > 5
            for slot in used:
                raw = tv["Nraw"][g, slot]
                n = np.where(raw == 0xFFFFFFFF, 0, raw & 0x1FFFF)
                assert int(n.sum()) == int(tv["ns"][g, slot])     # Ns[s] = sum_a Nsa[(s, a)]  (MCTS.py:139-144)


def _material_eval(x):
    """uniform priors, value = (own stacks - opponent stacks) / 49 from the planes: a smooth, meaningful evaluator (the
    hash stub's values are noise, under which ANY change of visiting order sends the search elsewhere)"""
    own = x[:, 0].reshape(x.shape[0], -1).sum(1)
    opp = x[:, 1].reshape(x.shape[0], -1).sum(1)
    return torch.ones(x.shape[0], 343, dtype=torch.float32, device=x.device), ((own - opp) / 49.0).to(torch.float32)


def test_visit_distribution_stays_close_and_net_calls_drop(backend):
    """32 mid-game positions, 200 simulations: the leaf-parallel search must land where the sequential one does --
    small total-variation distance between the root visit distributions, and no regret of its most-visited move when
    judged by the sequential search's own Q -- with about K times fewer, K times larger net calls."""
    from inflexion_b200.state import pack_states
    from oracle import inflexion_rules as R
    G, sims = 32, 200
    b, p, t, _ = R.random_reachable_states(400, seed=5)
    sel = np.nonzero((t > 8) & (t < 60))[0][:G]
    roots = pack_states(b[sel], p[sel], t[sel], 0, 343, tags=np.arange(G))
    res, calls = {}, {}
    for K in (1, 8):
        calls[K] = []

        def evaluate(x, K=K):
            calls[K].append(x.shape[0])
            return _material_eval(x)
        m = BatchedMCTS(G, sims, 1.0, seed=3, backend=backend, leaves_per_wave=K)
        m.reset(roots)
        m.search(evaluate)
        m.raise_on_faults()
        st = m.root_stats(want_q=True)
        res[K] = (st["counts"].cpu().numpy().astype(np.float64), st["q64"].cpu().numpy())
    (c1, q1), (c8, _) = res[1], res[8]
    assert (c1.sum(axis=1) == sims - 1).all() and (c8.sum(axis=1) == sims - 1).all()
    p1, p8 = c1 / c1.sum(axis=1, keepdims=True), c8 / c8.sum(axis=1, keepdims=True)
    assert (0.5 * np.abs(p1 - p8).sum(axis=1)).mean() < 0.15
    a1, a8 = c1.argmax(axis=1), c8.argmax(axis=1)
    regret = [q1[i, a1[i]] - q1[i, a8[i]] for i in range(G) if c1[i, a8[i]] > 0]
    assert len(regret) >= G - 2 and abs(float(np.mean(regret))) < 0.02
    assert len(calls[1]) == sims and set(calls[1]) == {G}
    assert len(calls[8]) <= sims // 8 + 12 and set(calls[8]) == {G * 8}


def test_object_api_leaves_per_wave(backend, monkeypatch):
    import inflexion_b200.backend as backend_mod
    from inflexion_b200.inflexion.InflexionGame import InflexionGame
    from inflexion_b200.MCTS import MCTS
    from inflexion_b200.utils import dotdict
    monkeypatch.setattr(backend_mod, "_DEFAULT", backend)
    rows = []

    class Net:
        def predict(self, planes):
            rows.append(1)
            pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None])
            return pi.numpy()[0], v.numpy()
    game = InflexionGame(7, max_turns=343, max_power=6)
    probs = MCTS(Net(), dotdict({"numMCTSSims": 30, "cpuct": 1, "leavesPerWave": 6}), seed=5, backend=backend).getActionProb(game, temp=1)
    assert abs(probs.sum() - 1.0) < 1e-12 and (probs[:294] == 0).all()      # empty board: only spawns are valid


def test_batched_arena_leaf_parallel(backend):
    """BatchedArena(leaves_per_wave=K): complete games, valid results, K times larger net batches."""
    from inflexion_b200.arena_batched import BatchedArena
    rows = []

    def evaluate(x):
        rows.append(x.shape[0])
        return stub_policy_value(x)
    res = BatchedArena(evaluate, 12, opponent="greedy", seed=4, max_turns=10, backend=backend, leaves_per_wave=4).playGames(6, shard=False)
    assert sum(res) == 6 and set(rows) == {24}


def test_selfplay_engine_leaf_parallel_emits_examples(backend):
    """SelfPlayEngine(leaves_per_wave=K): whole episodes, examples with the reference's shapes and value targets."""
    from inflexion_b200.selfplay import SelfPlayEngine
    eng = SelfPlayEngine(3, 16, lambda x: stub_policy_value(x), temp_threshold=3, seed=2, max_turns=6, backend=backend,
                         leaves_per_wave=4)
    moves = eng.play_episode()
    eng.mcts.raise_on_faults()
    planes, pi, v = eng.examples()
    assert moves == 7 and planes.shape == (3 * 7 * 36, 4, 7, 7) and np.allclose(pi.sum(axis=1), 1.0)
    assert set(np.round(np.abs(v), 4).tolist()) <= {1.0, 0.0001}


@pytest.mark.parametrize("K", [1, 4])
def test_captured_search_control_flow(K):
    """search_captured replays (net, wave) steps without looking at the device where it need not: with the graph
    replaced by its eager body (no CUDA here) it must leave the same trees as search / search_adaptive."""
    from hostsim import hostsim_backend
    be = hostsim_backend()
    G, sims = 5, 23
    out = []
    for captured in (False, True):
        m = BatchedMCTS(G, sims, 1.0, seed=9, backend=be, leaves_per_wave=K)
        m.reset(start_states(G, 12))
        ev = lambda x: stub_policy_value(x)  # noqa: E731
        if captured:
            class Step:
                def replay(self_inner):
                    m.feed(ev)
                    m.wave()
            m._captured_step = lambda evaluate, version: Step()
        stats = []
        for mv in range(4):
            if captured:
                m.search_captured(ev)
            else:
                m.search(ev)
            m.raise_on_faults()
            st = m.root_stats(want_q=True)
            stats.append((st["counts"].clone(), st["q32"].clone()))
            m.advance(m.sample_actions(0))
        out.append(stats)
    for (c0, q0), (c1, q1) in zip(*out):
        assert torch.equal(c0, c1) and torch.equal(q0.view(torch.int32), q1.view(torch.int32))
        assert int(c0.sum(1).min()) >= sims - 1
