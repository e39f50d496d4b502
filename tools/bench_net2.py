"""Dev tool: FoldedInference (current) at batch 4096 under a CUDA graph: ms per forward, TFLOP/s, error vs fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import inflexion_b200
from inflexion_b200.inflexion.pytorch.InflexionNNet import InflexionNNet, FoldedInference
from inflexion_b200.utils import dotdict

class Shape:
    board_shape, max_actions = (7, 7), 343
dev = torch.device("cuda", 0)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
net = InflexionNNet(Shape(), dotdict({"num_channels": 512, "dropout": 0.3})).to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.zeros(B, 4, 7, 7, device=dev)
x[:, 0] = (torch.rand(B, 7, 7, device=dev) < 0.25).float(); x[:, 1] = (torch.rand(B, 7, 7, device=dev) < 0.2).float() * (1 - x[:, 0])
x[:, 2] = 100.0; x[:, 3] = 1.0
FLOP = 404_287_488 * B
f = FoldedInference(net)
with torch.no_grad():
    lp, v = net(x[:512])
pi, vv = f(x[:512])
print("gemm_conv1", f.gemm_conv1, "max|dpi|", float((pi - torch.exp(lp)).abs().max()), "max|dv|", float((vv - v.reshape(-1)).abs().max()))
out = torch.empty(B, 344, dtype=torch.float16, device=dev)
for name, fn in (("head_into", lambda: f.head_into(x, out)), ("__call__", lambda: f(x))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50): g.replay()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 50
    print(f"{name:12s} B={B} {ms:7.3f} ms  {FLOP/ms/1e9:7.1f} TFLOP/s", flush=True)
