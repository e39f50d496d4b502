"""Soak run (GPU box): complete self-play episodes -- G games x 344 moves x S sims with the default net -- through
SelfPlayEngine.play_episode: no faults over whole games, measured games/s, example emission at the end."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import inflexion_b200
from inflexion_b200.backend import Backend
from inflexion_b200.inflexion.InflexionGame import InflexionGame
from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
from inflexion_b200.selfplay import SelfPlayEngine
from inflexion_b200.state import STATE_DTYPE

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
be = Backend.cuda(0)
torch.manual_seed(0); torch.backends.cudnn.benchmark = True
nnet = NNetWrapper(InflexionGame(7, max_turns=343, max_power=6), device=be.device)
eng = SelfPlayEngine(G, S, nnet.predict_head, fused_head=True, temp_threshold=30, seed=7, max_turns=343, backend=be)
eng.play_move(); torch.cuda.synchronize()                  # graph capture + warm-up (move 1 of the episode)
t0 = time.time()
moves = eng.play_episode()
eng.mcts.raise_on_faults()
torch.cuda.synchronize()
dt = time.time() - t0
st = eng.mcts.status().cpu().numpy()
c = eng.mcts.counters()
final = eng.mcts.roots().cpu().numpy().view(STATE_DTYPE).reshape(-1)
print(f"soak: {G} games x {moves} moves x {S} sims in {dt:.1f}s (+ first move): {G / dt:.1f} games/s, "
      f"{c['sims'] / dt:.3e} sims/s; outcomes draw/won/lost for the side to move: "
      f"{(final['outcome'] == 1).sum()}/{(final['outcome'] == 2).sum()}/{(final['outcome'] == 3).sum()}, unfinished {(final['outcome'] == 0).sum()}; "
      f"V/sim {c['node_visits'] / c['sims']:.2f}, terminal leaves {c['terminals'] / c['sims']:.3f}/sim; "
      f"live nodes max {st[:, 5].max()}, free nodes min {st[:, 6].min()}; faults {int((st[:, 2] != 0).sum())}")
t1 = time.time()
planes, pi, v = eng.examples(games=range(8))
print(f"examples of 8 games: {len(v)} (36 x {len(v) // 36 // 8} moves x 8) in {time.time() - t1:.2f}s; "
      f"pi rows sum to 1: {bool(np.allclose(pi.sum(axis=1), 1.0))}; v in {sorted(set(np.round(v, 4).tolist()))[:6]}")
