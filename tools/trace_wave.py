"""Per-warp timing trace of the tree wave kernel (dev tool): where do the microseconds of one wave go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import inflexion_b200
from inflexion_b200.backend import Backend
from inflexion_b200.selfplay import SelfPlayEngine
from inflexion_b200.stubnet import stub_policy_value

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
be = Backend.cuda(0)
eng = SelfPlayEngine(games, 100, lambda x: stub_policy_value(x), seed=2023, record=False, use_graph=False, backend=be)
for _ in range(4):
    eng.play_move()
dbg = torch.zeros((games, 4), dtype=torch.int64, device=be.device)
be.check(be.dll.ifx_mcts_set_debug(eng.mcts._h, dbg.data_ptr()))
for w in range(60):
    eng._wave_and_net()
    if w in (20, 40, 59):
        torch.cuda.synchronize()
        d = dbg.cpu().numpy()
        t0 = d[:, 0].min()
        start, end, depth, probes = (d[:, 0] - t0) / 1e3, (d[:, 1] - t0) / 1e3, d[:, 2], d[:, 3]
        life = end - start
        print(f"wave {w}: kernel span {end.max():.1f} us; warp start p50 {np.median(start):.1f} max {start.max():.1f}; "
              f"lifetime mean {life.mean():.1f} p50 {np.median(life):.1f} p90 {np.percentile(life, 90):.1f} max {life.max():.1f}")
        for dd in sorted(set(depth.tolist())):
            sel = depth == dd
            print(f"   depth {dd}: n={sel.sum():5d} life mean {life[sel].mean():6.1f} max {life[sel].max():6.1f} probes {probes[sel].mean():.2f}")
be.check(be.dll.ifx_mcts_set_debug(eng.mcts._h, None))
