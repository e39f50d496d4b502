"""Dev tool: time the folded inference graph variants of the policy/value net on one GPU (batch 4096)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import inflexion_b200
from inflexion_b200.inflexion.pytorch.InflexionNNet import InflexionNNet, FoldedInference, _fold
from inflexion_b200.utils import dotdict

class Shape:
    board_shape, max_actions = (7, 7), 343
dev = torch.device("cuda", 0)
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
net = InflexionNNet(Shape(), dotdict({"num_channels": 512, "dropout": 0.3})).to(dev).eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = torch.zeros(B, 4, 7, 7, device=dev)
x[:, 0] = (torch.rand(B, 7, 7, device=dev) < 0.25).float(); x[:, 2] = 100.0; x[:, 3] = 1.0
FLOP = 404_287_488 * B

def timeit(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"{name:40s} {ms:7.3f} ms  {FLOP/ms/1e9:7.1f} TFLOP/s", flush=True)
    return out

with torch.no_grad():
    lp, v = net(x[:256])
    ref_pi, ref_v = torch.exp(lp), v.reshape(-1)
for dt in (torch.bfloat16, torch.float16):
    f = FoldedInference(net, dt)
    pi, vv = f(x[:256])
    print(dt, "max|dpi|", float((pi - ref_pi).abs().max()), "max|dv|", float((vv - ref_v).abs().max()))
    timeit(lambda: f(x), f"FoldedInference {dt}")

# variant: cudnn fused conv+bias+relu
class Fused:
    def __init__(self, net, dt):
        self.dt = dt
        w, b = _fold(net.conv1.weight, net.conv1.bias, net.bn1)
        self.w1, self.b1 = w.contiguous(memory_format=torch.channels_last), b
        self.convs = []
        for conv, bn, pad in ((net.conv2, net.bn2, 1), (net.conv3, net.bn3, 0), (net.conv4, net.bn4, 0)):
            w, b = _fold(conv.weight, conv.bias, bn)
            self.convs.append((w.to(dt).contiguous(memory_format=torch.channels_last), b.to(dt), pad))
        w, b = _fold(net.fc1.weight, net.fc1.bias, net.fc_bn1)
        w = w.view(-1, 512, 3, 3).permute(0, 2, 3, 1).reshape(-1, 4608)
        self.fc1 = (w.to(dt).contiguous(), b.to(dt))
        w, b = _fold(net.fc2.weight, net.fc2.bias, net.fc_bn2)
        self.fc2 = (w.to(dt).contiguous(), b.to(dt))
        wh = torch.cat([net.fc3.weight, net.fc4.weight]).detach().float(); bh = torch.cat([net.fc3.bias, net.fc4.bias]).detach().float()
        self.head = (wh.to(dt).contiguous(), bh.to(dt))
    @torch.no_grad()
    def __call__(self, x, first16=False):
        x = x.contiguous(memory_format=torch.channels_last)
        if first16:
            s = torch.cudnn_convolution_relu(x.to(self.dt), self.w1.to(self.dt), self.b1.to(self.dt), [1, 1], [1, 1], [1, 1], 1)
        else:
            s = torch.cudnn_convolution_relu(x, self.w1, self.b1, [1, 1], [1, 1], [1, 1], 1).to(self.dt)
        for w, b, pad in self.convs:
            s = torch.cudnn_convolution_relu(s, w, b, [1, 1], [pad, pad], [1, 1], 1)
        s = s.permute(0, 2, 3, 1).reshape(s.shape[0], -1)
        s = F.relu(F.linear(s, *self.fc1))
        s = F.relu(F.linear(s, *self.fc2))
        out = F.linear(s, *self.head).float()
        return torch.exp(F.log_softmax(out[:, :343], dim=1)), torch.tanh(out[:, 343])
for dt in (torch.bfloat16, torch.float16):
    try:
        f = Fused(net, dt)
        pi, vv = f(x[:256])
        print(dt, "fused max|dpi|", float((pi - ref_pi).abs().max()), "max|dv|", float((vv - ref_v).abs().max()))
        timeit(lambda: f(x), f"cudnn_convolution_relu {dt}")
    except Exception as ex:
        print("fused variant failed:", repr(ex)[:300])
# per-layer timing of the fp16 folded graph
f = FoldedInference(net, torch.float16)
xs = x.contiguous(memory_format=torch.channels_last)
def layer_times():
    s = F.relu(F.conv2d(xs, f.w1, f.b1, padding=1)).to(f.dtype)
    outs = [s]
    for w, b, pad in f.convs:
        s = F.relu(F.conv2d(s, w, b, padding=pad)); outs.append(s)
    return outs
outs = layer_times()
timeit(lambda: F.relu(F.conv2d(xs, f.w1, f.b1, padding=1)).to(f.dtype), "conv1 fp32 + relu + cast")
for i, (w, b, pad) in enumerate(f.convs):
    inp = outs[i]
    timeit(lambda: F.conv2d(inp, w, b, padding=pad), f"conv{i+2} only")
    timeit(lambda: F.relu(F.conv2d(inp, w, b, padding=pad)), f"conv{i+2} + relu")
flat = outs[3].permute(0, 2, 3, 1).reshape(B, -1)
timeit(lambda: F.relu(F.linear(flat, *f.fc1)), "fc1 + relu")
