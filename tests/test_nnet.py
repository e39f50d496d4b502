"""The policy/value net wrapper: reference API, checkpoint format, folded inference numerics, DDP training."""
import os
import sys

import numpy as np
import pytest
import torch

from inflexion_b200.inflexion.pytorch import NNet as nnet_mod
from inflexion_b200.inflexion.pytorch.InflexionNNet import FoldedInference, InflexionNNet
from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
from inflexion_b200.utils import dotdict


class Shape:
    board_shape, max_actions = (7, 7), 343


def random_planes(n, seed=0):
    rng = np.random.default_rng(seed)
    own = rng.random((n, 49)) < 0.25
    opp = (rng.random((n, 49)) < 0.25) & ~own
    turn = rng.integers(0, 344, n)
    spawn = rng.integers(0, 2, n)
    x = np.stack([own, opp, np.broadcast_to(turn[:, None], (n, 49)), np.broadcast_to(spawn[:, None], (n, 49))], axis=1)
    return x.reshape(n, 4, 7, 7).astype(np.float32)


def randomize_bn(net, seed=1):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def test_architecture_matches_reference_parameter_names():
    net = InflexionNNet(Shape(), dotdict({"num_channels": 512, "dropout": 0.3}))
    assert sum(p.numel() for p in net.parameters()) == 12_526_424
    keys = set(net.state_dict())
    for k in ("conv1.weight", "conv4.bias", "bn1.running_mean", "bn4.num_batches_tracked", "fc1.weight", "fc_bn1.weight",
              "fc_bn2.running_var", "fc3.weight", "fc4.bias"):
        assert k in keys
    assert net.fc1.weight.shape == (1024, 4608) and net.fc3.weight.shape == (343, 512)
    if os.path.isdir("/root/reference/inflexion"):
        sys.path.insert(0, "/root/reference")
        from inflexion.pytorch.InflexionNNet import InflexionNNet as RefNet

        class G:
            max_actions = 343
            def to_planes(self):
                return np.zeros((4, 7, 7))
        ref = RefNet(G(), dotdict({"num_channels": 512, "dropout": 0.3}))
        assert {k: v.shape for k, v in ref.state_dict().items()} == {k: v.shape for k, v in net.state_dict().items()}
        net.load_state_dict(ref.state_dict())                                  # reference checkpoints load
        net.eval(); ref.eval()
        x = torch.from_numpy(random_planes(5))
        with torch.no_grad():
            a, b = net(x), ref(x)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_folded_inference_equals_eval_forward_fp32():
    torch.manual_seed(0)
    net = InflexionNNet(Shape(), dotdict({"num_channels": 512, "dropout": 0.3})).eval()
    randomize_bn(net)
    x = torch.from_numpy(random_planes(16))
    with torch.no_grad():
        lp, v = net(x)
    pi, vv = FoldedInference(net, torch.float32)(x)
    assert pi.shape == (16, 343) and vv.shape == (16,)
    assert torch.allclose(pi, torch.exp(lp), atol=2e-5) and torch.allclose(vv, v.reshape(-1), atol=2e-5)


def test_wrapper_api_and_checkpoint_roundtrip(tmp_path, monkeypatch):
    monkeypatch.setitem(nnet_mod.args, "cuda", False)
    torch.manual_seed(1)
    w = NNetWrapper(Shape(), device="cpu", infer_dtype=torch.float32)
    board = random_planes(1)[0].astype(np.int64)
    pi, v = w.predict(board)
    assert pi.dtype == np.float32 and pi.shape == (343,) and v.dtype == np.float32 and v.shape == (1,)
    assert abs(pi.sum() - 1) < 1e-5 and -1 <= v[0] <= 1
    pb, vb = w.predict_batch(torch.from_numpy(board[None].astype(np.float32)))
    assert np.allclose(pb[0].numpy(), pi, atol=2e-5) and abs(float(vb[0]) - v[0]) < 2e-5
    w.save_checkpoint(str(tmp_path), "x.pth.tar")
    ck = torch.load(os.path.join(tmp_path, "x.pth.tar"))
    assert list(ck) == ["state_dict"]                                          # NNet.py:109-111
    w2 = NNetWrapper(Shape(), device="cpu", infer_dtype=torch.float32)
    w2.load_checkpoint(str(tmp_path), "x.pth.tar")
    assert np.array_equal(w2.predict(board)[0], pi)
    with pytest.raises(FileNotFoundError):
        w2.load_checkpoint(str(tmp_path), "missing.pth.tar")


def test_train_reduces_loss(monkeypatch):
    monkeypatch.setitem(nnet_mod.args, "cuda", False)
    monkeypatch.setitem(nnet_mod.args, "epochs", 3)
    monkeypatch.setitem(nnet_mod.args, "batch_size", 32)
    monkeypatch.setitem(nnet_mod.args, "num_channels", 16)
    torch.manual_seed(0); np.random.seed(0)
    w = NNetWrapper(Shape(), device="cpu", infer_dtype=torch.float32)
    planes = random_planes(64).astype(np.int64)
    pis = np.zeros((64, 343)); pis[np.arange(64), np.arange(64) % 343] = 1
    examples = [(planes[i], pis[i].tolist(), float((-1) ** i)) for i in range(64)]
    hist = w.train(examples, verbose=False)
    assert len(hist) == 3 and hist[-1][0] < hist[0][0]
    assert w._folded is None


@pytest.mark.gpu
def test_16bit_folded_inference_within_tolerance_on_gpu():
    """north_star: with the real net, priors and values agree within 1e-3 of the fp32 module.  The default
    inference dtype (fp16, fp32 accumulation) must meet it; bf16 is measured too and meets it for the priors only
    (the unnormalised turn plane, values up to 343, makes the bf16 value head drift by up to ~1e-2)."""
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    net = InflexionNNet(Shape(), dotdict({"num_channels": 512, "dropout": 0.3})).to(dev).eval()
    x = torch.from_numpy(random_planes(2048, seed=3)).to(dev)
    with torch.no_grad():
        lp, v = net(x)
    tol = 1e-3
    pi, vv = FoldedInference(net)(x)                                # default dtype
    assert FoldedInference(net).dtype == torch.float16
    assert float((pi - torch.exp(lp)).abs().max()) < tol, float((pi - torch.exp(lp)).abs().max())
    assert float((vv - v.reshape(-1)).abs().max()) < tol, float((vv - v.reshape(-1)).abs().max())
    pi, vv = FoldedInference(net, torch.bfloat16)(x)
    assert float((pi - torch.exp(lp)).abs().max()) < tol
    assert float((vv - v.reshape(-1)).abs().max()) < 3e-2
    # with non-trivial BN statistics the folded graph must still track the module
    randomize_bn(net)
    with torch.no_grad():
        lp, v = net(x)
    pi, vv = FoldedInference(net)(x)
    assert float((pi - torch.exp(lp)).abs().max()) < tol and float((vv - v.reshape(-1)).abs().max()) < 5e-3


@pytest.mark.gpu
def test_head_into_equals_call_on_gpu():
    """predict_head hands the raw binary16 head output to the tree kernel; with softmax/tanh applied it must equal
    the ordinary predict_batch output."""
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    net = InflexionNNet(Shape(), dotdict({"num_channels": 512, "dropout": 0.3})).to(dev).eval()
    f = FoldedInference(net)
    x = torch.from_numpy(random_planes(512, seed=5)).to(dev)
    pi, v = f(x)
    out = torch.empty(512, 344, dtype=torch.float16, device=dev)
    f.head_into(x, out)
    pi2 = torch.exp(torch.log_softmax(out[:, :343].float(), dim=1))
    v2 = torch.tanh(out[:, 343].float())
    assert torch.allclose(pi, pi2, atol=1e-6) and torch.allclose(v, v2, atol=1e-6)


@pytest.mark.gpu
def test_im2col_kernel_matches_unfold_on_gpu():
    """ifx_net_im2col3x3 (conv1's input stage) == torch's unfold of the same planes, bit for bit, zero-padded to 48;
    302 games: the last block of five is partial."""
    from inflexion_b200 import _capi
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(random_planes(302, seed=9)).to(dev)
    out = torch.full((302 * 49 + 1, 48), 7.0, dtype=torch.float16, device=dev)
    dll = _capi.lib()
    _capi.check(dll.ifx_net_im2col3x3(x.data_ptr(), 302, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), dll)
    ref = torch.nn.functional.unfold(x, kernel_size=3, padding=1)            # [302, 36, 49], rows c*9 + ky*3 + kx
    ref = ref.permute(0, 2, 1).reshape(302 * 49, 36).to(torch.float16)
    assert torch.equal(out[:-1, :36], ref) and float(out[:-1, 36:].abs().max()) == 0.0
    assert bool((out[-1] == 7.0).all())                          # nothing written past the last row
