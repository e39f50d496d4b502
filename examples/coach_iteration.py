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
    nnet = nnet_mod.NNetWrapper(game)
    coach = Coach(game, nnet, dotdict({"numMCTSSims": a.sims, "cpuct": 1, "tempThreshold": 30, "numEps": a.games * world}))

    t0 = time.time()
    examples = coach.selfPlayBatch(a.games * world, seed=1, gather=a.gather, compact=a.compact)   # this rank's games (all with --gather)
    torch.cuda.synchronize()
    t1 = time.time()
    moves = a.max_turns + 1
    if rank == 0:
        print(f"self-play: {a.games * world} games x {moves} moves x {a.sims} sims on {world} GPU(s) in {t1 - t0:.1f}s "
              f"({a.games * world * moves * a.sims / (t1 - t0):.3e} sims/s incl. example expansion); "
              f"{len(examples)} examples on rank 0", flush=True)
    np.random.seed(rank)                                        # every rank draws its own batches
    hist = nnet.train(examples, verbose=False)                  # gradients all-reduced across ranks
    w = torch.cat([p.detach().reshape(-1) for p in nnet.nnet.parameters()])
    if world > 1:                                               # ranks must hold identical weights after training
        ref = w.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, w), "ranks diverged"
    if rank == 0:
        print("train: loss_pi %.3f -> %.3f, loss_v %.3f -> %.3f (weights identical on all ranks)"
              % (hist[0][0], hist[-1][0], hist[0][1], hist[-1][1]), flush=True)
    for opp in ("random", "greedy"):
        t2 = time.time()
        one, two, dr = BatchedArena(nnet.predict_head, a.sims, opponent=opp, seed=2, max_turns=a.max_turns,
                                    fused_head=True).playGames(a.arena)
        if rank == 0:
            print(f"arena vs {opp}: MCTS {one} / {opp} {two} / draws {dr}  ({a.arena} games in {time.time() - t2:.1f}s)", flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
