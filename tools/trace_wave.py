"""Per-warp timing trace of the tree wave kernel (dev tool): where do the microseconds of one wave go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import inflexion_b200
from inflexion_b200.backend import Backend
from inflexion_b200.selfplay import SelfPlayEngine
from inflexion_b200.stubnet import stub_policy_value

games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
net = sys.argv[2] if len(sys.argv) > 2 else "stub"
be = Backend.cuda(0)
if net.startswith("default"):
    from inflexion_b200.inflexion.InflexionGame import InflexionGame
    from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
    torch.manual_seed(0); torch.backends.cudnn.benchmark = True
    nn = NNetWrapper(InflexionGame(7, max_turns=343, max_power=6), device=be.device)
    fused = net == "default"                  # "default" = the bench configuration (binary16 head fused into the tree kernel)
    evaluate = nn.predict_head if fused else nn.predict_batch     # "default-exact": float32 pi / v, exact renormalisation
else:
    evaluate, fused = (lambda x: stub_policy_value(x)), False
eng = SelfPlayEngine(games, 100, evaluate, seed=2023, record=False, use_graph=False, backend=be, fused_head=fused)
for _ in range(4):
    eng.play_move()
dbg = torch.zeros((games, 8), dtype=torch.int64, device=be.device)
be.check(be.dll.ifx_mcts_set_debug(eng.mcts._h, dbg.data_ptr()))
for w in range(60):
    eng._wave_and_net()
    if w in (20, 40, 59):
        torch.cuda.synchronize()
        d = dbg.cpu().numpy()
        t0 = d[:, 0].min()
        start, end, depth, probes = (d[:, 0] - t0) / 1e3, (d[:, 5] - t0) / 1e3, d[:, 6], d[:, 7]
        ok = (d[:, 2] > 0) & (d[:, 4] > 0)
        ph = [(d[ok, i + 1] - d[ok, i]) / 1e3 for i in range(4)]
        print("   phases (us, mean): loads %.2f  expand %.2f  backup %.2f  first level %.2f  | rest %.2f" % (
            ph[0].mean(), ph[1].mean(), ph[2].mean(), ph[3].mean(), ((d[ok, 5] - d[ok, 4]) / 1e3).mean()))
        life = end - start
        print(f"wave {w}: kernel span {end.max():.1f} us; warp start p50 {np.median(start):.1f} max {start.max():.1f}; "
              f"lifetime mean {life.mean():.1f} p50 {np.median(life):.1f} p90 {np.percentile(life, 90):.1f} max {life.max():.1f}")
        for dd in sorted(set(depth.tolist())):
            sel = depth == dd
            print(f"   depth {dd}: n={sel.sum():5d} life mean {life[sel].mean():6.1f} max {life[sel].max():6.1f} probes {probes[sel].mean():.2f}")
be.check(be.dll.ifx_mcts_set_debug(eng.mcts._h, None))
