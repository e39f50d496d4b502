#!/usr/bin/env python
"""One full Coach iteration on the device (BASELINE config 5): self-play sharded over the ranks -> NNetWrapper.train
with the gradient all-reduce (NCCL under torchrun) -> BatchedArena against Random and Greedy, games sharded over the
ranks and only the W/L/D counters all-reduced.

    python examples/coach_iteration.py --games 32 --sims 25 --arena 20 --max-turns 40
    torchrun --nproc-per-node 8 examples/coach_iteration.py --games 64 --sims 200 --arena 100
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

import inflexion_b200  # noqa: F401
from inflexion_b200.arena_batched import BatchedArena
from inflexion_b200.Coach import Coach
from inflexion_b200.inflexion.InflexionGame import InflexionGame
from inflexion_b200.inflexion.pytorch import NNet as nnet_mod
from inflexion_b200.utils import dotdict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=32, help="self-play games per GPU")
    ap.add_argument("--sims", type=int, default=25)
    ap.add_argument("--max-turns", type=int, default=343)
    ap.add_argument("--arena", type=int, default=20)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--gather", action="store_true", help="all-gather the un-expanded self-play records (final example gather)")
    ap.add_argument("--local-bn", action="store_true",
                    help="data-parallel training with per-rank BatchNorm statistics (NNet args.dp_sync_bn = False)")
    ap.add_argument("--compact", action="store_true",
                    help="keep un-expanded records and expand examples per training batch (selfplay.ExampleSet): thousands of games")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)                                        # identical initial weights on every rank
    game = InflexionGame(7, max_turns=a.max_turns, max_power=6)
    nnet_mod.args["epochs"] = a.epochs
    if a.local_bn:
        nnet_mod.args["dp_sync_bn"] = False
    nnet = nnet_mod.NNetWrapper(game)
    coach = Coach(game, nnet, dotdict({"numMCTSSims": a.sims, "cpuct": 1, "tempThreshold": 30, "numEps": a.games * world}))

    t0 = time.time()
    examples = coach.selfPlayBatch(a.games * world, seed=1, gather=a.gather, compact=a.compact)   # this rank's games (all with --gather)
    torch.cuda.synchronize()
    t1 = time.time()
    if rank == 0:
        print(f"self-play: {a.games * world} games (<= {a.max_turns + 1} moves each) x {a.sims} sims on {world} GPU(s) in {t1 - t0:.1f}s "
              f"= {a.games * world / (t1 - t0):.1f} games/s incl. the example gather/expansion; {len(examples)} examples on rank 0",
              flush=True)
    n_local = torch.tensor([len(examples)], dtype=torch.int64, device="cuda")
    n_all = [torch.zeros_like(n_local) for _ in range(world)]
    if world > 1:
        dist.all_gather(n_all, n_local)
    else:
        n_all = [n_local]
    np.random.seed(rank)                                        # every rank draws its own batches
    torch.cuda.synchronize()
    t_tr = time.time()
    hist = nnet.train(examples, verbose=False)                  # gradients all-reduced across ranks (DDP buckets)
    torch.cuda.synchronize()
    t_tr = time.time() - t_tr
    w = torch.cat([t.detach().reshape(-1).float() for t in nnet.nnet.state_dict().values()])
    if world > 1:                                               # ranks must hold identical nets after training
        ref = w.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, w), "ranks diverged"
    steps = a.epochs * int(sum(int(x) for x in n_all) / nnet_mod.args.batch_size)
    if rank == 0:
        print(f"examples per rank: {[int(x) for x in n_all]} (games end at different moves, so the counts may differ)", flush=True)
        print("train" + (" (per-rank BatchNorm statistics)" if a.local_bn else "") + ": %d steps in %.1fs (%.2f ms/step, global batch %d = %d per rank); loss_pi %.3f -> %.3f, loss_v %.3f -> %.3f; "
              "parameters and BatchNorm buffers identical on all ranks"
              % (steps, t_tr, 1e3 * t_tr / max(steps, 1), nnet_mod.args.batch_size, nnet_mod.args.batch_size // world,
                 hist[0][0], hist[-1][0], hist[0][1], hist[-1][1]), flush=True)
    if world > 1:
        # how much of a training step is exposed all-reduce: the same step with and without gradient synchronisation
        from torch.nn.parallel import DistributedDataParallel as DDP
        ddp = DDP(nnet.nnet, device_ids=[local], broadcast_buffers=False, gradient_as_bucket_view=True, bucket_cap_mb=8)
        opt = torch.optim.Adam(nnet.nnet.parameters())
        per = nnet_mod.args.batch_size // world
        x = torch.randint(0, 2, (per, 4, 7, 7), device="cuda").float()
        tp = torch.full((per, 343), 1.0 / 343, device="cuda")
        tv = torch.zeros(per, device="cuda")
        nnet.nnet.train()

        def run(sync, n=30):
            torch.cuda.synchronize(); dist.barrier(); t = time.time()
            for _ in range(n):
                ctx = ddp.no_sync() if not sync else torch.autograd.profiler.record_function("sync")
                with ctx:
                    lp, v = ddp(x)
                    loss = nnet.loss_pi(tp, lp) + nnet.loss_v(tv, v)
                    opt.zero_grad(set_to_none=False)
                    loss.backward()
                opt.step()
            torch.cuda.synchronize(); dist.barrier()
            return (time.time() - t) / n
        run(True, 5); run(False, 5)
        t_sync, t_nosync = run(True), run(False)
        nnet.nnet.eval()
        nnet.weights_changed()
        if rank == 0:
            print("all-reduce share of a train step: %.2f ms with the bucketed all-reduce overlapped with backward vs %.2f ms "
                  "without gradient synchronisation -> %.1f %% exposed (50 MB of fp32 gradients per step over NVLink)"
                  % (1e3 * t_sync, 1e3 * t_nosync, 100 * max(t_sync - t_nosync, 0) / t_sync), flush=True)
    for opp in ("random", "greedy"):
        t2 = time.time()
        one, two, dr = BatchedArena(nnet.predict_head, a.sims, opponent=opp, seed=2, max_turns=a.max_turns,
                                    fused_head=True, capture=True).playGames(a.arena)
        if rank == 0:
            print(f"arena vs {opp}: MCTS {one} / {opp} {two} / draws {dr}  ({a.arena} games in {time.time() - t2:.1f}s)", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
