"""Leaf gather (north_star kernel 3): the net batch holds only the games that search in the current move.
ifx_mcts_live_rows + ifx_gather_pack_rows / ifx_gather_scatter_rows through the C ABI, and the packed search against
the dense one (identical visit counts and actions: the reference evaluates one leaf per real simulation, MCTS.py:89-93,
so which ROW of a batch a leaf travels in must not matter)."""
import numpy as np
import torch

from inflexion_b200.arena_batched import BatchedArena
from inflexion_b200.mcts_batched import BatchedMCTS
from inflexion_b200.selfplay import SelfPlayEngine
from inflexion_b200.stubnet import stub_policy_value
from test_mcts_kernel import start_states


def test_pack_and_scatter_rows_abi(backend):
    be, dev = backend, backend.device
    G, cap = 37, 16
    gen = torch.Generator().manual_seed(1)
    for cols, dtype in ((196, torch.float32), (344, torch.float16), (1, torch.float32), (343, torch.float32)):
        src = torch.randn(G, cols, generator=gen).to(dtype).to(dev)
        idx = torch.tensor([3, 5, 8, 13, 21, 34, 36] + [0] * (G - 7), dtype=torch.int32, device=dev)
        n = torch.tensor([7], dtype=torch.int32, device=dev)
        packed = torch.full((cap, cols), 9, dtype=dtype, device=dev)
        row_bytes = cols * src.element_size()
        be.check(be.dll.ifx_gather_pack_rows(src.data_ptr(), row_bytes, idx.data_ptr(), n.data_ptr(), cap, row_bytes,
                                             packed.data_ptr(), row_bytes, be.stream()))
        assert torch.equal(packed[:7], src[idx[:7].long()]) and bool((packed[7:] == 0).all())
        dst = torch.zeros_like(src)
        be.check(be.dll.ifx_gather_scatter_rows(packed.data_ptr(), row_bytes, idx.data_ptr(), n.data_ptr(), cap, row_bytes,
                                                dst.data_ptr(), row_bytes, be.stream()))
        want = torch.zeros_like(src)
        want[idx[:7].long()] = src[idx[:7].long()]
        assert torch.equal(dst, want)
    # bad arguments are refused, not executed
    assert be.dll.ifx_gather_pack_rows(src.data_ptr(), 2, idx.data_ptr(), n.data_ptr(), cap, 6, packed.data_ptr(), 6, be.stream()) == -1


def _fused_eval(nn_in, out):
    """a deterministic binary16 head from the planes (fused-head evaluators write [rows, 344] in place)"""
    pi, v = stub_policy_value(nn_in, "dyadic")
    out[:, :343] = (pi % 23 - 11).to(torch.float16)
    out[:, 343] = v.to(torch.float16)


def test_packed_search_equals_dense_search(backend):
    G = 70
    max_turns = np.where(np.arange(G) % 3 == 0, 343, 3)        # two thirds of the games end after 4 moves
    for fused in (False, True):
        outs = []
        for quantum in (0, 32):
            rows_seen = []

            def evaluate(x, out=None):
                rows_seen.append(x.shape[0])
                return _fused_eval(x, out) if fused else stub_policy_value(x)
            eng = SelfPlayEngine(G, 6, evaluate, temp_threshold=3, seed=5, max_turns=343, backend=backend, use_graph=False,
                                 fused_head=fused)
            eng.mcts.pack_quantum = quantum
            eng.mcts.reset(start_states(G, max_turns))
            for _ in range(7):
                eng.play_move()
            eng.mcts.raise_on_faults()
            outs.append((eng.ex_counts[:7].clone(), eng.ex_actions[:7].clone(), sorted(set(rows_seen)), eng.mcts.rows_evaluated))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert outs[0][2] == [G]                                   # dense: every evaluation on all rows
        assert outs[1][2] == [32, G]                               # packed: 24 live games -> one 32-row bucket
        assert outs[1][3] < 0.8 * outs[0][3]


def test_live_rows_lists_the_searching_games(backend):
    G = 40
    m = BatchedMCTS(G, 4, 1.0, seed=2, backend=backend, pack_quantum=8)
    roots = start_states(G, 343)
    roots["outcome"][::4] = 1                                   # every fourth game is already over
    m.reset(roots)
    assert m.plan_rows() == 32 and m.live_count == 30
    live = m._live_idx[:30].cpu().numpy()
    assert np.array_equal(live, np.array([g for g in range(G) if g % 4]))
    m.search(lambda x: stub_policy_value(x))                     # after the search nobody has simulations left
    assert m.plan_rows() == 8 and m.live_count == 0


def test_arena_packs_the_games_the_mcts_player_moves_in(backend):
    rows = []

    def evaluate(x):
        rows.append(x.shape[0])
        return stub_policy_value(x)
    arena = BatchedArena(evaluate, 5, opponent="greedy", seed=4, max_turns=10, backend=backend)
    one, two, draws = arena.playGames(70, shard=False)
    assert one + two + draws == 70
    # RED (the MCTS player) moves first in 36 of the 70 games: every half-ply searches about half of the games
    assert max(rows) <= 64 and min(rows) >= 32
    dense = BatchedArena(lambda x: stub_policy_value(x), 5, opponent="greedy", seed=4, max_turns=10, backend=backend)
    import inflexion_b200.mcts_batched as mb
    orig = mb.BatchedMCTS.plan_rows
    mb.BatchedMCTS.plan_rows = lambda self: self.n_games
    try:
        assert dense.playGames(70, shard=False) == (one, two, draws)
    finally:
        mb.BatchedMCTS.plan_rows = orig


def test_wave_parts_equal_the_whole_wave(backend):
    """ifx_mcts_wave_part: games are independent, so searching them as two parts (each with its own net call on its own
    rows) gives the same tables as one wave over all of them."""
    G, sims = 9, 12
    outs = []
    for parts in (None, ((0, 4), (4, 5))):
        m = BatchedMCTS(G, sims, 1.0, seed=6, backend=backend)
        m.reset(start_states(G, 343))
        for _ in range(2):
            for _ in range(m.waves_per_search - 1):
                if parts is None:
                    m.wave()
                    m.feed(lambda x: stub_policy_value(x))
                else:
                    for p in parts:
                        m.wave(p)
                        m.feed(lambda x: stub_policy_value(x), p)
            for p in (parts or (None,)):
                m.wave(p)
            m.raise_on_faults()
            acts = m.sample_actions(3)
            outs.append((m.root_stats()["counts"].clone(), acts.clone()))
            m.advance(acts)
    for a, b in zip(outs[:2], outs[2:]):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert m.be.dll.ifx_mcts_wave_part(m._h, 5, 5, m.nn_pi.data_ptr(), 352, m.nn_v.data_ptr(), m.nn_in.data_ptr(), None) == -1


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("net", ["stub", "default"])
def test_two_stream_capture_equals_single_stream(cuda_backend, net):
    """SelfPlayEngine(streams=2): the captured move runs (wave, net) of the two halves of the games on two streams.
    Same visit counts and actions as the single-stream capture (stub net: bit for bit; the CNN: the two halves go
    through the same kernels on their own rows, so equality is expected there too)."""
    from inflexion_b200.stubnet import stub_kernel_evaluator
    be = cuda_backend
    outs = []
    for streams in (1, 2):
        if net == "stub":
            eng = SelfPlayEngine(512, 20, None, seed=3, backend=be, streams=streams, auto_restart=True)
            eng.evaluate = stub_kernel_evaluator(eng.mcts)
        else:
            from inflexion_b200.inflexion.InflexionGame import InflexionGame
            from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
            torch.manual_seed(0)
            nnet = NNetWrapper(InflexionGame(7, max_turns=343, max_power=6), device=be.device)
            eng = SelfPlayEngine(512, 20, nnet.predict_head, seed=3, backend=be, streams=streams, auto_restart=True, fused_head=True)
        for _ in range(3):
            eng.play_move()
        eng.mcts.raise_on_faults()
        torch.cuda.synchronize()
        outs.append((eng.ex_counts[:3].clone(), eng.ex_actions[:3].clone()))
    if net == "stub":
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    else:                                                      # (batch-size dependent kernel choices may move a prior by an ulp)
        same = (outs[0][0] == outs[1][0]).all(dim=2).float().mean()
        assert float(same) > 0.97


@pytest.mark.gpu
def test_packed_cuda_graphs_per_bucket_equal_the_dense_eager_search(cuda_backend):
    """2 304 games whose episodes end at different moves: the captured move is re-used per net batch size (one graph per
    bucket of 256 rows, the live count read on the device), and every bucket must give what the dense, un-captured search
    gives -- visit counts and actions of every game and move."""
    from inflexion_b200.stubnet import stub_kernel_evaluator
    be = cuda_backend
    G, sims, moves = 2304, 12, 9
    max_turns = 2 + (np.arange(G) % 7)                         # episodes of 3..9 moves: the live set shrinks move by move
    outs, buckets = [], []
    for graph in (False, True):
        eng = SelfPlayEngine(G, sims, None, temp_threshold=3, seed=8, max_turns=343, backend=be, use_graph=graph)
        eng.evaluate = stub_kernel_evaluator(eng.mcts)
        eng.mcts.pack_quantum = 256 if graph else 0
        eng.mcts.reset(start_states(G, max_turns))
        for _ in range(moves):
            eng.play_move()
            buckets.append(eng.mcts.rows_active)
        eng.mcts.raise_on_faults()
        torch.cuda.synchronize()
        outs.append((eng.ex_counts[:moves].clone(), eng.ex_actions[:moves].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert len(set(buckets[moves:])) >= 4 and min(buckets[moves:]) <= 512      # several bucket sizes were captured and replayed
