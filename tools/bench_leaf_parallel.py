"""Dev tool (GPU box): what the leaf-parallel mode (virtual loss, leaves_per_wave = K) buys where the sequential search is
net-latency-bound -- a single MCTS object (Arena / pit.py through the object API) and small sets of games -- with the
default CNN.  Prints moves/s (one move = numMCTSSims simulations per game) and net calls per move."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import inflexion_b200
from inflexion_b200.backend import Backend
from inflexion_b200.inflexion.InflexionGame import InflexionGame
from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
from inflexion_b200.mcts_batched import BatchedMCTS
from inflexion_b200.state import pack_states

be = Backend.cuda(0)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
nnet = NNetWrapper(InflexionGame(7, max_turns=343, max_power=6), device=be.device)
sims = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for G in (1, 16, 256):
    for K in (1, 4, 8, 16):
        m = BatchedMCTS(G, sims, 1.0, seed=1, backend=be, fused_head=True, leaves_per_wave=K)
        m.reset(pack_states(np.zeros((G, 49), dtype=np.int64), 1, 0, 0, 343, tags=np.arange(G)))
        calls = [0]

        def evaluate(x, out):
            calls[0] += 1
            return nnet.predict_head(x, out)
        for mv in range(2):                                   # warm-up moves
            (m.search_adaptive if K == 1 else m.search)(evaluate)
            m.advance(m.sample_actions(0))
        torch.cuda.synchronize()
        calls[0] = 0
        moves = 6
        t0 = time.time()
        for mv in range(moves):
            (m.search_adaptive if K == 1 else m.search)(evaluate)
            m.advance(m.sample_actions(0))
        torch.cuda.synchronize()
        dt = time.time() - t0
        m.raise_on_faults()
        print(f"games {G:4d}  leaves_per_wave {K:2d}: {moves / dt:7.2f} moves/s = {G * sims * moves / dt:10.3e} sims/s, "
              f"{calls[0] / moves:6.1f} net calls per move (batch {G * K})", flush=True)
