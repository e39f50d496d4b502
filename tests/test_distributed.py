"""N>1 paths on CPU with the gloo backend (world_size 2): games shard by rank with no data-path collective,
and NNetWrapper.train averages gradients across ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _selfplay_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import inflexion_b200  # noqa: F401
    from hostsim import hostsim_backend
    from inflexion_b200.selfplay import SelfPlayEngine
    from inflexion_b200.stubnet import stub_policy_value
    per_rank = 3
    eng = SelfPlayEngine(per_rank, 10, lambda x: stub_policy_value(x), temp_threshold=3, seed=5, max_turns=5,
                         game_id_base=rank * per_rank, backend=hostsim_backend())
    moves = eng.play_episode()
    acts = eng.ex_actions[:6].clone()
    acts[moves:] = -1
    # the only communication: gathering the (small) results, as the final example gather would
    gathered = [torch.zeros_like(acts) for _ in range(world)]
    dist.all_gather(gathered, acts)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), torch.cat(gathered, dim=1).numpy())
    dist.destroy_process_group()


def test_sharded_selfplay_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_selfplay_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sharded = np.load(os.path.join(tmp_path, "sharded.npy"))
    from hostsim import hostsim_backend
    from inflexion_b200.selfplay import SelfPlayEngine
    from inflexion_b200.stubnet import stub_policy_value
    eng = SelfPlayEngine(6, 10, lambda x: stub_policy_value(x), temp_threshold=3, seed=5, max_turns=5,
                         backend=hostsim_backend())
    moves = eng.play_episode()
    single = eng.ex_actions[:6].clone()
    single[moves:] = -1
    assert np.array_equal(sharded, single.numpy())          # game i plays the same game on whichever rank owns it


def _train_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import inflexion_b200  # noqa: F401
    from inflexion_b200.inflexion.pytorch import NNet as nnet_mod
    from test_nnet import Shape, random_planes
    nnet_mod.args.update(cuda=False, epochs=2, batch_size=16, num_channels=8)
    torch.manual_seed(rank)                                  # DIFFERENT initial weights per rank: train() must broadcast rank 0's
    w = nnet_mod.NNetWrapper(Shape(), device="cpu", infer_dtype=torch.float32)
    np.random.seed(100 + rank)                               # different batches per rank
    n = 48 + 40 * rank                                       # rank-local example counts differ (games end at different
    planes = random_planes(n, seed=rank).astype(np.int64)    # moves): the ranks must still agree on the number of steps
    pis = np.full((n, 343), 1.0 / 343)
    examples = [(planes[i], pis[i].tolist(), float((-1) ** (i + rank))) for i in range(n)]
    w.train(examples, verbose=False)
    # parameters AND buffers (BatchNorm running statistics): the whole state a checkpoint holds
    flat = torch.cat([t.detach().reshape(-1).float() for t in w.nnet.state_dict().values()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    # a rank without examples: EVERY rank must raise (decided collectively), nobody may hang in an all-reduce
    raised = 0
    try:
        w.train([] if rank == 1 else examples, verbose=False)
    except ValueError:
        raised = 1
    flags = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([raised]))
    if rank == 0:
        torch.save(gathered + [torch.cat(flags)], os.path.join(out_dir, "weights.pt"))
    dist.destroy_process_group()


def test_data_parallel_training_keeps_ranks_in_sync(tmp_path):
    port = _free_port()
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w0, w1, raised = torch.load(os.path.join(tmp_path, "weights.pt"))
    assert torch.equal(w0, w1)                               # identical nets although the ranks started from different ones
    assert raised.tolist() == [1, 1]                         # the empty-rank error is raised on every rank
    assert float(w0.abs().sum()) > 0


def _coach_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import inflexion_b200  # noqa: F401
    from hostsim import hostsim_backend
    from inflexion_b200.arena_batched import BatchedArena
    from inflexion_b200.selfplay import SelfPlayEngine, build_examples
    from inflexion_b200.stubnet import stub_policy_value
    be = hostsim_backend()
    # sharded Arena: 6 games over 2 ranks, only the W/L/D counters are all-reduced
    res = BatchedArena(lambda x: stub_policy_value(x), 6, opponent="greedy", seed=4, max_turns=8, backend=be).playGames(6)
    # final example gather: un-expanded records of both ranks, expanded after the gather
    per_rank = 2
    eng = SelfPlayEngine(per_rank, 8, lambda x: stub_policy_value(x), temp_threshold=3, seed=5, max_turns=5,
                         game_id_base=rank * per_rank, backend=be)
    while eng.move < eng.max_moves:
        eng.play_move()
    planes, pi, v = build_examples(be, eng.gather_records())
    if rank == 0:
        np.savez(os.path.join(out_dir, "coach.npz"), res=np.asarray(res), planes=planes, pi=pi, v=v)
    dist.destroy_process_group()


def test_sharded_arena_and_example_gather(tmp_path):
    port = _free_port()
    mp.spawn(_coach_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(tmp_path, "coach.npz"))
    from hostsim import hostsim_backend
    from inflexion_b200.arena_batched import BatchedArena
    from inflexion_b200.selfplay import SelfPlayEngine
    from inflexion_b200.stubnet import stub_policy_value
    be = hostsim_backend()
    single = BatchedArena(lambda x: stub_policy_value(x), 6, opponent="greedy", seed=4, max_turns=8, backend=be).playGames(6)
    assert tuple(got["res"]) == single and sum(single) == 6
    eng = SelfPlayEngine(4, 8, lambda x: stub_policy_value(x), temp_threshold=3, seed=5, max_turns=5, backend=be)
    while eng.move < eng.max_moves:
        eng.play_move()
    planes, pi, v = eng.examples()
    assert len(v) > 0 and np.array_equal(got["planes"], planes) and np.array_equal(got["pi"], pi) and np.array_equal(got["v"], v)


def _learn_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import inflexion_b200  # noqa: F401
    import inflexion_b200.backend as backend_mod
    from hostsim import hostsim_backend
    from inflexion_b200.Coach import Coach
    from inflexion_b200.inflexion.InflexionGame import InflexionGame
    from inflexion_b200.inflexion.pytorch import NNet as nnet_mod
    from inflexion_b200.utils import dotdict
    backend_mod._DEFAULT = hostsim_backend()                 # test seam: the kernels' logic-check build on the CPU
    nnet_mod.args.update(cuda=False, epochs=1, batch_size=32, num_channels=8)
    torch.manual_seed(rank)                                  # DIFFERENT initial weights per rank: train() must broadcast rank 0's
    np.random.seed(3)
    game = InflexionGame(7, max_turns=6, max_power=6)
    nnet = nnet_mod.NNetWrapper(game, device="cpu", infer_dtype=torch.float32)
    args = dotdict({"numIters": 1, "numEps": 4, "tempThreshold": 3, "numMCTSSims": 6, "cpuct": 1, "maxlenOfQueue": 200000,
                    "numItersForTrainExamplesHistory": 2, "checkpoint": os.path.join(out_dir, "ckpt"), "arenaCompare": 4,
                    "pitInterval": 1})                        # pit (sharded Arena vs Random and Greedy) in this iteration too
    coach = Coach(game, nnet, args)
    coach.learn()                                            # self-play (2 games per rank) -> train (all-reduce) -> checkpoint
    n_local = sum(len(d) for d in coach.trainExamplesHistory)
    # a second iteration in compact mode: un-expanded records, examples expanded per training batch
    args["compactExamples"] = True
    compact_coach = Coach(game, nnet, args)
    compact_coach.learn()
    assert len(compact_coach.trainExamplesHistory) == 1 and len(compact_coach.trainExamplesHistory[0]) % 36 == 0
    flat = torch.cat([p.detach().reshape(-1) for p in nnet.nnet.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n_local]))
    if rank == 0:
        torch.save({"weights": gathered, "counts": [int(c) for c in counts]}, os.path.join(out_dir, "learn.pt"))
    dist.destroy_process_group()


def test_coach_learn_iteration_under_torch_distributed(tmp_path):
    """Coach.learn with an initialised process group: games shard by rank, example counts differ per rank, training
    stays in lock-step (same weights everywhere), rank 0 writes the checkpoint, every rank its own example shard."""
    port = _free_port()
    mp.spawn(_learn_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(tmp_path, "learn.pt"))
    assert torch.equal(got["weights"][0], got["weights"][1])
    assert all(c > 0 and c % 36 == 0 for c in got["counts"])
    ck = os.path.join(tmp_path, "ckpt")
    assert os.path.isfile(os.path.join(ck, "temp.pth.tar"))
    assert os.path.isfile(os.path.join(ck, "checkpoint_0.pth.tar.examples.rank0"))
    assert os.path.isfile(os.path.join(ck, "checkpoint_0.pth.tar.examples.rank1"))
