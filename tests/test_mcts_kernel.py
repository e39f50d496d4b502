"""The tree kernels (through the C ABI) against the golden episodes frozen from the live reference and against
the oracle: root visit counts, Q bit patterns and kinds, Ns, sampled actions -- all bit exact."""
import os

import numpy as np
import pytest
import torch

from inflexion_b200.mcts_batched import BatchedMCTS
from inflexion_b200.state import STATE_DTYPE, pack_states, unpack_boards
from inflexion_b200.stubnet import stub_policy_value
from oracle.inflexion_mcts import OracleMCTS, PhiloxDraws, sample_action, self_play_moves
from oracle.inflexion_rules import OracleGame

EPISODES = ["a_float_25", "b_dyadic_50", "c_short_full", "d_short_full"]


def start_states(n, max_turns, tags=None):
    return pack_states(np.zeros((n, 49), dtype=np.int64), 1, 0, 0, max_turns,
                       tags=np.arange(n) if tags is None else tags)


def stub_eval(mode):
    return lambda nn_in: stub_policy_value(nn_in, mode)


def stub_predict(mode):
    def predict(planes):
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None], mode)
        return pi.numpy()[0], v.numpy()[0]
    return predict


@pytest.mark.parametrize("name", EPISODES)
def test_golden_episode(backend, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"mcts_{name}.npz"))
    cpuct = int(g["cpuct"]) if bool(g["cpuct_is_int"]) else float(g["cpuct"])
    mode, sims, thr = str(g["mode"]), int(g["sims"]), int(g["temp_threshold"])
    m = BatchedMCTS(1, sims, cpuct, seed=int(g["seed"]), backend=backend)
    m.reset(start_states(1, int(g["max_turns"])))
    for i in range(len(g["action"])):
        m.search(stub_eval(mode))
        m.raise_on_faults()
        st = m.root_stats(want_q=True)
        counts = st["counts"].cpu().numpy()[0]
        assert np.array_equal(counts, g["counts"][i]), f"visit counts differ at move {i}"
        seen = counts > 0
        assert np.array_equal(st["qkind"].cpu().numpy()[0][seen], g["q_is32"][i][seen])
        assert st["q64"].cpu().numpy()[0][seen].tobytes() == g["q"][i][seen].tobytes(), f"Q bits differ at move {i}"
        assert int(st["ns"][0]) == g["ns"][i]
        root = m.roots().cpu().numpy().view(STATE_DTYPE).reshape(-1)[0]
        assert np.array_equal(unpack_boards(np.array([root]))[0], g["board"][i]) and root["turn"] == g["turn"][i]
        a = m.sample_actions(thr)
        assert int(a[0]) == g["action"][i], f"sampled action differs at move {i}"
        m.advance(a)
    status = m.status().cpu().numpy()[0]
    assert status[2] == 0
    if "ex_v" in g.files:
        assert status[4] != 0          # the episode ended exactly where the reference's did
    assert m.counters()["expansions"] == int(g["net_calls"])


def test_many_games_match_oracle(backend):
    """G independent trees (different RNG streams, different max_turns) against one oracle run each."""
    G, sims, cpuct, seed, moves = 6, 20, 1.5, 77, 7
    max_turns = np.array([343, 343, 6, 4, 343, 5])
    m = BatchedMCTS(G, sims, cpuct, seed=seed, backend=backend)
    m.reset(start_states(G, max_turns, tags=np.arange(G) + 100))
    oracles = [self_play_moves(stub_predict("float"), sims, cpuct, PhiloxDraws(seed, game=100 + i), temp_threshold=3,
                               max_moves=moves, max_turns=int(max_turns[i])) for i in range(G)]
    alive = [True] * G
    for mv in range(moves):
        m.search(stub_eval("float"))
        m.raise_on_faults()
        counts = m.root_stats()["counts"].cpu().numpy()
        acts = m.sample_actions(3)
        a_np = acts.cpu().numpy()
        for i in range(G):
            if not alive[i]:
                assert a_np[i] == -1
                continue
            rec = next(oracles[i])
            assert np.array_equal(counts[i], rec["counts"]), (mv, i)
            assert a_np[i] == rec["action"], (mv, i)
            if rec["game"].outcome != 0:
                alive[i] = False
        m.advance(acts)
    assert not all(alive) and any(alive)


def test_full_tables_match_oracle(backend):
    """Every live node (not just the root): P rows, N, Q and Ns against the oracle's dictionaries."""
    sims, cpuct, seed = 30, 1, 5
    m = BatchedMCTS(1, sims, cpuct, seed=seed, backend=backend)
    m.reset(start_states(1, 8))
    draws = PhiloxDraws(seed)
    game = OracleGame(max_turns=8)
    orc = OracleMCTS(stub_predict("float"), sims, cpuct, draws)
    for mv in range(7):
        orc.move = mv
        pi = orc.get_action_prob(game, temp=1)
        m.search(stub_eval("float"))
        tv = m.table_views()
        keys = tv["key"][0]
        used = np.nonzero(keys[:, 0] != np.uint64(2 ** 64 - 1))[0]
        # oracle nodes still reachable (turn > root turn, or the root)
        live = {k: v for k, v in orc.P.items()
                if np.frombuffer(k, dtype=np.int64)[2 * 49] > game.turn or k == game.key()}
        assert len(used) == len(live)
        for slot in used:
            k0, k1 = int(keys[slot, 0]), int(keys[slot, 1])
            own = np.array([(k0 >> c) & 1 for c in range(49)], dtype=np.int64)
            opp = np.array([(k1 >> c) & 1 for c in range(49)], dtype=np.int64)
            planes = np.stack([own, opp, np.full(49, k0 >> 49), np.full(49, (k1 >> 49) & 1)]).astype(np.int64)
            key = planes.reshape(4, 7, 7).tobytes()
            assert key in live
            assert tv["P"][0, slot].tobytes() == orc.P[key].astype(np.float32).tobytes()
            assert tv["ns"][0, slot] == orc.Ns[key]
            for a in range(343):
                raw = int(tv["Nraw"][0, slot, a])
                n = 0 if raw == 0xFFFFFFFF else raw & 0x1FFFF
                assert (raw == 0xFFFFFFFF) == (orc.V[key][a] == 0)
                assert n == orc.N.get((key, a), 0)
                if n:
                    q = orc.Q[(key, a)]
                    is64 = bool(raw >> 31)
                    assert is64 == (not q.is32)
                    got = tv["Q64"][0, slot, a] if is64 else np.float64(tv["Q32"][0, slot, a])
                    assert np.float64(got).tobytes() == np.float64(q.x).tobytes()
        a = sample_action(pi, draws.uniform(mv))
        acts = m.sample_actions(1000)
        assert int(acts[0]) == a
        m.advance(acts)
        game = game.to_next_state(a)
        if game.outcome != 0:
            break


def test_uniform_fallback_when_all_valid_priors_are_zero(backend):
    """MCTS.py:98-107: a policy that is zero on every valid move -> uniform over the valid moves."""
    m = BatchedMCTS(2, 5, 1.0, seed=1, backend=backend)
    m.reset(start_states(2, 343))

    def zero_eval(nn_in):
        return torch.zeros(2, 343, device=nn_in.device), torch.full((2,), 0.25, device=nn_in.device)
    m.search(zero_eval)
    m.raise_on_faults()
    tv = m.table_views()
    counts = m.root_stats()["counts"].cpu().numpy()
    # root prior: 49 spawn actions, each 1/49 in float32; first sim expands the root, the other 4 visit children
    assert counts.sum(axis=1).tolist() == [4, 4]
    used = np.nonzero(tv["key"][0][:, 0] != np.uint64(2 ** 64 - 1))[0]
    roots = [s for s in used if (int(tv["key"][0][s, 0]) >> 49) == 0]
    p = tv["P"][0, roots[0]]
    assert np.all(p[294:] == np.float32(1.0) / np.float32(49.0)) and np.all(p[:294] == 0)
    # unseen edges tie at u = cpuct * P * sqrt(Ns + 1e-8): lowest index wins until visited
    assert counts[0][294:298].tolist() == [1, 1, 1, 1]


def test_pool_exhaustion_is_reported(backend):
    m = BatchedMCTS(1, 40, 1.0, seed=1, nodes_per_game=32, backend=backend)
    m.reset(start_states(1, 343))
    m.search(stub_eval("float"))
    with pytest.raises(Exception, match="pool"):
        m.raise_on_faults()


def test_fused_head_softmax_matches_torch(backend):
    """fused-head mode: the kernel's softmax / tanh of the net's binary16 head output against torch's
    exp(log_softmax) / tanh (NNet.py:94, InflexionNNet.py:54), then masked renormalisation -- float32 tolerance."""
    G = 5
    m = BatchedMCTS(G, 3, 1.0, seed=3, backend=backend, fused_head=True)
    m.reset(start_states(G, 343))
    gen = torch.Generator().manual_seed(0)
    head = (torch.randn(G, 344, generator=gen) * 3).to(torch.float16)

    def evaluate(nn_in, out):
        out.copy_(head.to(out.device))
    m.search(evaluate)
    m.raise_on_faults()
    tv = m.table_views()
    x = head.float()
    probs = torch.exp(torch.log_softmax(x[:, :343], dim=1))
    valid = torch.zeros(343); valid[294:] = 1                      # empty board: the 49 spawn actions
    want = (probs * valid) / (probs * valid).sum(dim=1, keepdim=True)
    for g in range(G):
        keys = tv["key"][g]
        root = [s for s in np.nonzero(keys[:, 0] != np.uint64(2 ** 64 - 1))[0] if (int(keys[s, 0]) >> 49) == 0][0]
        got = torch.from_numpy(tv["P"][g, root].copy())
        assert torch.allclose(got, want[g], rtol=2e-5, atol=1e-8)
    # the root was expanded with v = tanh(head[:, 343]); its children were then visited twice
    counts = m.root_stats()["counts"].cpu().numpy()
    assert counts.sum(axis=1).tolist() == [2] * G


def test_dense_policy_rows_give_the_same_search(backend):
    """nn_pi with a row stride of 343 floats (unaligned rows, scalar loads) == padded rows (128-bit loads)."""
    out = []
    for pad in (True, False):
        m = BatchedMCTS(3, 15, 1.0, seed=8, backend=backend, pad_policy_rows=pad)
        m.reset(start_states(3, 343))
        m.search(stub_eval("float"))
        m.raise_on_faults()
        out.append(m.root_stats(want_q=True))
    assert m.nn_pi.stride(0) == 343
    for k in ("counts", "q32", "ns"):
        assert torch.equal(out[0][k].cpu(), out[1][k].cpu())


def test_terminal_budget_does_not_change_the_search(backend):
    """terminal_sims_per_wave only decides in which wave a simulation runs: with games that end after 4-6 moves
    (most simulations reach a terminal state) num_sims+1 waves still complete every move, with identical tables."""
    G, sims, moves = 5, 24, 5
    max_turns = np.array([4, 5, 6, 343, 5])
    ref = None
    for budget in (0, 1, 2):
        m = BatchedMCTS(G, sims, 1.0, seed=21, backend=backend, terminal_sims_per_wave=budget)
        m.reset(start_states(G, max_turns, tags=np.arange(G)))
        out = []
        for _ in range(moves):
            m.search(stub_eval("float"))
            m.raise_on_faults()
            st = m.status().cpu().numpy()
            over = st[:, 4] != 0
            assert np.all((st[:, 0] == sims) | over) and not st[:, 1].any()
            stats = m.root_stats(want_q=True)
            acts = m.sample_actions(2)
            out.append((stats["counts"].cpu().numpy().copy(), stats["q32"].cpu().numpy().copy(), acts.cpu().numpy().copy()))
            m.advance(acts)
        if ref is None:
            ref = out
        else:
            for (c0, q0, a0), (c1, q1, a1) in zip(ref, out):
                assert np.array_equal(c0, c1) and q0.tobytes() == q1.tobytes() and np.array_equal(a0, a1)


@pytest.mark.parametrize("mode", ["float", "dyadic"])
def test_stub_kernel_equals_torch_stub(backend, mode):
    """csrc/ifx_stub.cu is stubnet.stub_policy_value bit for bit (it is the evaluator of tree-only throughput runs)."""
    from inflexion_b200.stubnet import stub_kernel_evaluator
    G = 37
    m = BatchedMCTS(G, 4, 1.0, seed=2, backend=backend)
    rng = np.random.default_rng(5)
    x = np.zeros((G, 4, 49), dtype=np.float32)
    own = rng.random((G, 49)) < 0.3
    x[:, 0] = own; x[:, 1] = (rng.random((G, 49)) < 0.3) & ~own
    x[:, 2] = rng.integers(0, 344, G)[:, None]; x[:, 3] = rng.integers(0, 2, G)[:, None]
    nn_in = torch.from_numpy(x.reshape(G, 4, 7, 7)).to(backend.device)
    pi, v = stub_kernel_evaluator(m, mode)(nn_in)
    pi_ref, v_ref = stub_policy_value(nn_in, mode)
    assert torch.equal(pi.cpu(), pi_ref.cpu()) and torch.equal(v.cpu(), v_ref.cpu())


def test_optional_dirichlet_root_noise(backend):
    """Not in the reference (default off).  With dirichlet_eps > 0 the root priors become (1-eps) P + eps eta,
    eta ~ Dir(alpha) over the valid actions, once per move: eta >= 0, sums to 1, has the Dirichlet mean and variance
    across games; the search stays deterministic in the seed; eps = 0 leaves everything bit-identical."""
    G, sims, alpha, eps = 96, 6, 0.3, 0.25

    def run(**kw):
        m = BatchedMCTS(G, sims, 1.0, seed=13, backend=backend, **kw)
        m.reset(start_states(G, 343))
        m.search(stub_eval("float"))
        m.raise_on_faults()
        tv = m.table_views()
        roots = []
        for g in range(G):                                   # the root is the node whose turn is 0
            used = np.nonzero(tv["key"][g, :, 0] != np.uint64(2 ** 64 - 1))[0]
            slot = [s for s in used if int(tv["key"][g, s, 0]) >> 49 == 0][0]
            roots.append(tv["P"][g, slot].copy())
        return np.stack(roots), m.root_stats()["counts"].cpu().numpy()
    p0, c0 = run()
    p0b, c0b = run(dirichlet_alpha=alpha, dirichlet_eps=0.0)
    assert p0.tobytes() == p0b.tobytes() and np.array_equal(c0, c0b)
    p1, c1 = run(dirichlet_alpha=alpha, dirichlet_eps=eps)
    p1b, c1b = run(dirichlet_alpha=alpha, dirichlet_eps=eps)
    assert p1.tobytes() == p1b.tobytes() and np.array_equal(c1, c1b)        # deterministic in the seed
    assert not np.array_equal(c0, c1)                                        # and it does change the search
    valid = p0 > 0                                                           # start position: the 49 spawn moves
    nv = valid.sum(axis=1)
    assert np.all(nv == 49)
    eta = (p1.astype(np.float64) - (1 - eps) * p0) / eps
    assert np.all(eta[valid] > -1e-5) and np.all(eta[~valid] == 0)
    assert np.allclose(eta.sum(axis=1), 1.0, atol=1e-4) and np.allclose(p1.sum(axis=1), 1.0, atol=1e-4)
    e = eta[valid].reshape(G, 49)
    var_theory = (1 / 49) * (1 - 1 / 49) / (49 * alpha + 1)
    assert abs(e.mean() - 1 / 49) < 1e-6                                     # rows sum to one
    assert 0.75 * var_theory < e.var() < 1.3 * var_theory, (e.var(), var_theory)
    # different games draw different noise
    assert len({e[g].tobytes() for g in range(G)}) == G


def test_small_tables_probe_chains_and_slot_reuse(backend):
    """A tight node pool (64 slots, 128 hash entries: load factor up to 0.5, probe chains longer than one 8-entry
    group, slots reclaimed and reused every move) must give the same search as the default-sized tables."""
    G, sims, moves = 4, 20, 6
    out = []
    for nodes in (None, 64):
        m = BatchedMCTS(G, sims, 1.0, seed=31, backend=backend, nodes_per_game=nodes)
        m.reset(start_states(G, 343))
        rec = []
        for _ in range(moves):
            m.search(stub_eval("float"))
            m.raise_on_faults()
            s = m.root_stats(want_q=True)
            acts = m.sample_actions(3)
            rec.append((s["counts"].cpu().numpy().copy(), s["q32"].cpu().numpy().copy(), acts.cpu().numpy().copy()))
            m.advance(acts)
        out.append(rec)
    assert m.cfg.nodes_per_game == 64 and m.cfg.hash_per_game == 128
    for (c0, q0, a0), (c1, q1, a1) in zip(*out):
        assert np.array_equal(c0, c1) and q0.tobytes() == q1.tobytes() and np.array_equal(a0, a1)


def test_zero_sized_calls_are_noops(backend):
    dll = backend.dll
    assert dll.ifx_stub_policy_value(None, 0, None, 0, None, 343, None, None) == 0
    assert dll.ifx_stub_policy_value(None, 1, None, 0, None, 343, None, None) == -1      # IFX_E_ARG
    assert dll.ifx_net_im2col3x3(None, 0, None, None) == 0


def sharp_policy_value(planes: torch.Tensor):
    """A sharply peaked deterministic evaluator: the actions are ranked by a hash of the position and the action ranked
    r gets the (unnormalised) prior 2^(-4r), value 0 -- the best-ranked VALID move of a node takes >= 15/16 of its prior
    mass after the masked renormalisation, so the principal variation grows by about one edge per simulation, which is
    what a confident trained net does to the search.  Powers of two and integer sorting only: the same bits everywhere."""
    from inflexion_b200.stubnet import _coef, _mix, _MASK31
    x = planes.reshape(planes.shape[0], 196).to(torch.int64)
    h = _mix((x * _coef(x.device)).sum(dim=1) & _MASK31)
    a = torch.arange(343, device=x.device, dtype=torch.int64)
    z = _mix((h[:, None] + a[None, :] * 0x9E3779B) & _MASK31)
    rank = torch.argsort(torch.argsort(z * 512 + a[None, :], dim=1), dim=1)         # unique keys: no tie handling involved
    pi = torch.ldexp(torch.ones_like(rank, dtype=torch.float32), -torch.clamp(4 * rank, max=124).to(torch.int32))
    return pi, torch.zeros(planes.shape[0], dtype=torch.float32, device=x.device)


def test_paths_deeper_than_one_register_chunk_match_oracle(backend):
    """the reference recurses without limit; with a confident policy the search path passes 32 edges within
    a few moves at 100 simulations.  Paths beyond the 32 levels a warp holds in registers spill to the workspace; visit
    counts, sampled actions and the number of live nodes must still equal the oracle's, with the default node pool."""
    sims, cpuct, seed, moves = 100, 1.0, 11, 4

    def predict(planes):
        pi, v = sharp_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None])
        return pi.numpy()[0], v.numpy()[0]
    m = BatchedMCTS(2, sims, cpuct, seed=seed, backend=backend)
    assert m.cfg.max_depth == 352
    m.reset(start_states(2, 343))
    deepest = [0]
    for gi in range(2):
        it = self_play_moves(predict, sims, cpuct, PhiloxDraws(seed, game=gi), temp_threshold=2, max_moves=moves)
        if gi == 0:
            m_search = OracleMCTS.search

            def tracked(self, game, _d=[0]):
                _d[0] += 1
                deepest[0] = max(deepest[0], _d[0] - 1)
                try:
                    return m_search(self, game)
                finally:
                    _d[0] -= 1
            OracleMCTS.search = tracked
        try:
            recs = [next(it) for _ in range(moves)]
        finally:
            OracleMCTS.search = m_search
        if gi == 0:
            want0 = recs
        else:
            want1 = recs
    assert deepest[0] > 40, deepest                     # well past one register chunk (and past round 1's cap of 32)
    for mv in range(moves):
        m.search(sharp_policy_value)
        m.raise_on_faults()
        counts = m.root_stats()["counts"].cpu().numpy()
        acts = m.sample_actions(2)
        for gi, want in enumerate((want0, want1)):
            assert np.array_equal(counts[gi], want[mv]["counts"]), (mv, gi)
            assert int(acts[gi]) == want[mv]["action"], (mv, gi)
        live = m.status().cpu().numpy()[:, 5]
        orc, game = want0[mv]["mcts"], want0[mv]["game"]
        m.advance(acts)
    # live nodes after the last advance: the oracle's nodes with turn > root turn, or the root itself
    live = m.status().cpu().numpy()[0, 5]
    want_live = sum(1 for k in orc.P if np.frombuffer(k, dtype=np.int64)[2 * 49] > game.turn or k == game.key())
    assert live == want_live


def test_max_depth_fault_and_limits(backend):
    """max_depth below the path the search needs -> IFX_F_DEPTH, loudly; num_sims >= 2^17 is refused."""
    m = BatchedMCTS(1, 60, 1.0, seed=11, max_depth=8, backend=backend)
    m.reset(start_states(1, 343))
    m.search(sharp_policy_value)
    with pytest.raises(Exception, match="deeper"):
        m.raise_on_faults()
    with pytest.raises(Exception):
        BatchedMCTS(1, 1 << 17, 1.0, backend=backend)
