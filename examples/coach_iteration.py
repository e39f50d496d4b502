#!/usr/bin/env python
"""One full Coach iteration on the device (BASELINE config 5): batched self-play -> NNetWrapper.train (gradient
all-reduce over NCCL when launched with torchrun) -> BatchedArena against Random and Greedy.

    python examples/coach_iteration.py --games 256 --sims 25 --arena 20 --epochs 1
    torchrun --nproc-per-node 8 examples/coach_iteration.py --games 4096 --sims 200 --arena 100
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
from inflexion_b200.inflexion.InflexionGame import InflexionGame
from inflexion_b200.inflexion.pytorch import NNet as nnet_mod
from inflexion_b200.selfplay import SelfPlayEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=256, help="self-play games per GPU")
    ap.add_argument("--sims", type=int, default=25)
    ap.add_argument("--max-turns", type=int, default=343)
    ap.add_argument("--arena", type=int, default=20)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--max-moves", type=int, default=None, help="cut self-play after this many moves (smoke runs)")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)                                        # identical initial weights on every rank
    game = InflexionGame(7, max_turns=a.max_turns, max_power=6)
    nnet_mod.args["epochs"] = a.epochs
    nnet = nnet_mod.NNetWrapper(game)

    t0 = time.time()
    eng = SelfPlayEngine(a.games, a.sims, nnet.predict_head, fused_head=True, temp_threshold=30, seed=1,
                         max_turns=a.max_turns, game_id_base=rank * a.games)
    moves = eng.play_episode(a.max_moves)
    eng.mcts.raise_on_faults()
    torch.cuda.synchronize()
    t1 = time.time()
    planes, pi, v = eng.examples(games=range(min(a.games, 8)))  # expand a few games' examples for the training step
    examples = [(planes[i], pi[i].tolist(), float(v[i])) for i in range(len(v))]
    if rank == 0:
        print(f"self-play: {a.games * world} games x {moves} moves x {a.sims} sims in {t1 - t0:.1f}s "
              f"({a.games * world * moves * a.sims / (t1 - t0):.3e} sims/s); {len(examples)} examples from {min(a.games, 8)} games")
    if len(examples) >= nnet_mod.args.batch_size:
        np.random.seed(rank)
        hist = nnet.train(examples, verbose=False)              # rank-local examples, all-reduced gradients
        if rank == 0:
            print("train: loss_pi %.3f -> %.3f, loss_v %.3f -> %.3f" % (hist[0][0], hist[-1][0], hist[0][1], hist[-1][1]))
    if rank == 0:
        for opp in ("random", "greedy"):
            one, two, dr = BatchedArena(nnet.predict_batch, a.sims, opponent=opp, seed=2, max_turns=a.max_turns).playGames(a.arena)
            print(f"arena vs {opp}: MCTS {one} / {opp} {two} / draws {dr}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
