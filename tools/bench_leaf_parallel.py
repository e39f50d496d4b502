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
        line = f"games {G:4d}  leaves_per_wave {K:2d}:"
        for captured in (False, True):
            m = BatchedMCTS(G, sims, 1.0, seed=1, backend=be, fused_head=True, leaves_per_wave=K)
            m.reset(pack_states(np.zeros((G, 49), dtype=np.int64), 1, 0, 0, 343, tags=np.arange(G)))

            def search():
                if captured:
                    m.search_captured(nnet.predict_head)
                else:
                    (m.search_adaptive if K == 1 else m.search)(nnet.predict_head)
            for mv in range(2):                                   # warm-up moves (and the capture)
                search()
                m.advance(m.sample_actions(0))
            torch.cuda.synchronize()
            e0, moves = m.evals_run, 6
            t0 = time.time()
            for mv in range(moves):
                search()
                m.advance(m.sample_actions(0))
            torch.cuda.synchronize()
            dt = time.time() - t0
            m.raise_on_faults()
            line += (f"  {'captured step' if captured else 'eager'}: {moves / dt:7.2f} moves/s = {G * sims * moves / dt:9.3e} sims/s, "
                     f"{(m.evals_run - e0) / moves:6.1f} net calls/move;")
        print(line + f"  (net batch {G * K})", flush=True)

# the drop-in object API (MCTS.getActionProb, one game): captured by default, args.captureSearch = False = eager
from inflexion_b200.MCTS import MCTS
from inflexion_b200.utils import dotdict
for K in (1, 16):
    for captured in (False, True):
        args = dotdict({"numMCTSSims": sims, "cpuct": 1.0, "captureSearch": captured, "leavesPerWave": K})
        mcts, game = MCTS(nnet, args, seed=3, backend=be), InflexionGame(7, max_turns=343, max_power=6)
        for mv in range(2):
            game = game.to_next_state(int(np.argmax(mcts.getActionProb(game, temp=0))))
        torch.cuda.synchronize()
        t0, moves = time.time(), 8
        for mv in range(moves):
            game = game.to_next_state(int(np.argmax(mcts.getActionProb(game, temp=0))))
        dt = time.time() - t0
        print(f"MCTS.getActionProb (object API, 1 game, {sims} sims, leavesPerWave {K}, "
              f"{'captured step' if captured else 'eager'}): {moves / dt:7.2f} moves/s = {sims * moves / dt:9.3e} sims/s", flush=True)
