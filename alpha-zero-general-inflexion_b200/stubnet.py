"""Deterministic policy/value stub used for bit-exact parity runs and for tree-only throughput runs.

The parity contract (BASELINE.json north_star) is "visit counts bit-exact against the reference's MCTS
when both use a fixed deterministic net stub".  A stub that is to give the SAME float32 numbers on the
CPU (reference / oracle side) and on the GPU (kernel side) must avoid every operation whose rounding is
implementation defined, so it uses only int64 arithmetic, an int->float32 conversion and one IEEE float32
division -- all of which torch evaluates identically on both devices.

``stub_policy_value(planes)`` takes the network input batch ``[B,4,7,7]`` (any real dtype; it is the
reference's ``to_planes()`` stack, ``inflexion/InflexionGame.py:84-91``) and returns ``(pi f32[B,343],
v f32[B])`` with the dtypes/shapes of ``NNetWrapper.predict`` (``inflexion/pytorch/NNet.py:94``).
"""
from __future__ import annotations

import torch

_MASK31 = (1 << 31) - 1
_COEF = {}      # per device (a host->device copy is not allowed while a CUDA graph is being captured)


def _coef(device):
    key = str(device)
    if key not in _COEF:
        if "cpu" not in _COEF:
            g = torch.Generator().manual_seed(20231018)
            _COEF["cpu"] = torch.randint(1, 1 << 20, (196,), generator=g, dtype=torch.int64) | 1
        _COEF[key] = _COEF["cpu"].to(device)
    return _COEF[key]


def _divisor(device):
    """float32(1000003) as a DEVICE tensor: dividing by a Python scalar lets torch's CUDA kernel multiply by the
    reciprocal instead, which rounds differently from the CPU's true division."""
    key = "div:" + str(device)
    if key not in _COEF:
        _COEF[key] = torch.tensor(1000003.0, dtype=torch.float32).to(device)
    return _COEF[key]


def _mix(h):
    h = (h ^ (h >> 15)) & _MASK31
    h = (h * 0x2C1B3C6D) & _MASK31
    h = (h ^ (h >> 12)) & _MASK31
    h = (h * 0x297A2D39) & _MASK31
    return (h ^ (h >> 15)) & _MASK31


def stub_kernel_evaluator(mcts, mode: str = "float"):
    """The same stub as ONE kernel of libifx_b200 (csrc/ifx_stub.cu), writing straight into the search's net-output
    buffers: evaluate(nn_in) -> (mcts.nn_pi, mcts.nn_v), so BatchedMCTS.feed copies nothing.  For tree-and-rules-only
    throughput runs, where the torch version below (a dozen elementwise launches) would dominate the step."""
    be = mcts.be
    coef = _coef(be.device)
    assert not mcts.fused_head

    def evaluate(nn_in):
        # rows of the search's own input buffer (all of it, or one half on its own stream) are answered in the same rows
        # of nn_pi / nn_v; any other batch (a packed one) in rows 0..n-1, which the caller then scatters
        n = nn_in.shape[0]
        off = nn_in.data_ptr() - mcts.nn_in.data_ptr()
        r0 = off // 784 if (0 <= off < mcts.n_games * 784 and off % 784 == 0) else 0
        pi, v = mcts.nn_pi[r0:r0 + n], mcts.nn_v[r0:r0 + n]
        be.check(be.dll.ifx_stub_policy_value(nn_in.data_ptr(), n, coef.data_ptr(), int(mode == "dyadic"),
                                              pi.data_ptr(), mcts.nn_pi.stride(0), v.data_ptr(), be.stream()))
        return (mcts.nn_pi, mcts.nn_v) if n == mcts.n_games else (pi, v)
    return evaluate


def stub_policy_value(planes: torch.Tensor, mode: str = "float"):
    """mode "dyadic": integer priors in [1,256] (every partial sum exact in float32, any order);
    mode "float": priors k/1000003 (inexact, so the float32 summation ORDER is observable)."""
    x = planes.reshape(planes.shape[0], 196).to(torch.int64)
    h = _mix((x * _coef(x.device)).sum(dim=1) & _MASK31)                  # [B]
    a = torch.arange(343, device=x.device, dtype=torch.int64)
    z = _mix((h[:, None] + a[None, :] * 0x9E3779B) & _MASK31)             # [B,343]
    sharp = (z & 0xFF) + 1                                                # 1..256
    if mode == "dyadic":
        pi = sharp.to(torch.float32)
    else:
        # a few large priors make the trees several plies deep, like a trained net's would
        peaked = torch.where((z >> 8) % 11 == 0, sharp * 64, sharp)
        pi = peaked.to(torch.float32) / _divisor(x.device)                # one IEEE float32 division per element
    hv = _mix(h ^ 0x5BD1E995)
    v = ((hv & 0xFFFF) - 32768).to(torch.float32) / 32768.0               # dyadic in [-1, 1)
    return pi, v
