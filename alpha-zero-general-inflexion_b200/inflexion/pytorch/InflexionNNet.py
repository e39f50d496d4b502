"""The policy/value CNN (reference architecture: inflexion/pytorch/InflexionNNet.py:11-54).

Parameter names and shapes are those of the reference so its ``*.pth.tar`` checkpoints load unchanged:
conv1..4 (3x3; the last two unpadded: 7->5->3) each with BatchNorm2d, fc1 4608->1024 and fc2 1024->512 with
BatchNorm1d and dropout, heads fc3 512->343 (log-softmax) and fc4 512->1 (tanh).  12 526 424 parameters,
404 287 488 FLOP per evaluation.

``forward`` is the training-time graph.  Self-play does not call it: ``FoldedInference`` below is the
inference form used by ``NNetWrapper.predict_batch`` -- BatchNorm folded into the preceding weights,
channels-last 16-bit activations on the tensor cores with conv+bias+ReLU as one fused cuDNN call
(cuDNN / cuBLAS; this is the only dense contraction on the path and stays in PyTorch by design).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class InflexionNNet(nn.Module):
    def __init__(self, game, args):
        super().__init__()
        self.depth, self.board_x, self.board_y = 4, *game.board_shape
        self.action_size = game.max_actions
        self.args = args
        ch = args.num_channels
        self.conv1 = nn.Conv2d(self.depth, ch, 3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(ch, ch, 3, stride=1, padding=1)
        self.conv3 = nn.Conv2d(ch, ch, 3, stride=1)
        self.conv4 = nn.Conv2d(ch, ch, 3, stride=1)
        self.bn1, self.bn2, self.bn3, self.bn4 = (nn.BatchNorm2d(ch) for _ in range(4))
        self.flat = ch * (self.board_x - 4) * (self.board_y - 4)
        self.fc1 = nn.Linear(self.flat, 1024)
        self.fc_bn1 = nn.BatchNorm1d(1024)
        self.fc2 = nn.Linear(1024, 512)
        self.fc_bn2 = nn.BatchNorm1d(512)
        self.fc3 = nn.Linear(512, self.action_size)
        self.fc4 = nn.Linear(512, 1)

    def forward(self, s):
        s = s.view(-1, self.depth, self.board_x, self.board_y)
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3), (self.conv4, self.bn4)):
            s = F.relu(bn(conv(s)))
        s = s.reshape(-1, self.flat)
        s = F.dropout(F.relu(self.fc_bn1(self.fc1(s))), p=self.args.dropout, training=self.training)
        s = F.dropout(F.relu(self.fc_bn2(self.fc2(s))), p=self.args.dropout, training=self.training)
        return F.log_softmax(self.fc3(s), dim=1), torch.tanh(self.fc4(s))


def _fold(weight, bias, bn):
    """y = bn(conv(x)) in eval mode == conv'(x) with w' = w * g/sqrt(var+eps), b' = (b - mean) * g/sqrt(var+eps) + beta."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = weight.detach().float() * scale.reshape(-1, *([1] * (weight.dim() - 1)))
    b = (bias.detach().float() - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return w, b


def _capi_check(rc, dll):
    from ..._capi import check
    check(rc, dll)


def _conv_relu(x, w, b, pad):
    """conv + bias + ReLU in one cuDNN call (fused epilogue) on CUDA; the three-op form elsewhere."""
    if x.is_cuda:
        return torch.cudnn_convolution_relu(x, w, b, [1, 1], [pad, pad], [1, 1], 1)
    return F.relu(F.conv2d(x, w, b, padding=pad))


class FoldedInference:
    """Eval-mode forward of an ``InflexionNNet`` with BN folded in; call with float32 [B,4,7,7] on the device.

    dtype float16 (default): every layer on the 16-bit tensor cores with fp32 accumulation; the input planes
    (0/1 masks and the turn counter <= 343) are exact in fp16.  dtype bfloat16: conv1 stays in fp32 (bf16 cannot
    hold integers above 256).  Measured against the fp32 module on a B200 (random-init weights,
    tests/test_nnet.py): fp16 max |dpi| 2e-6, max |dv| 3e-4; bf16 max |dpi| 1e-5, max |dv| 2e-3..8e-3 -- only fp16
    meets the 1e-3 value tolerance, which is why it is the default; both run at the same tensor-core rate.
    """

    def __init__(self, net: InflexionNNet, dtype=torch.float16):
        self.dtype = dtype
        dev = next(net.parameters()).device
        self.first_dtype = dtype if dtype == torch.float16 else torch.float32
        w1, b1 = _fold(net.conv1.weight, net.conv1.bias, net.bn1)
        self.w1 = w1.to(self.first_dtype).contiguous(memory_format=torch.channels_last)
        self.b1 = b1.to(self.first_dtype)
        # conv1 as a GEMM over im2col rows written by libifx_b200 (csrc/ifx_net.cu): [B*49, 48] x [48, 512], taps
        # c*9 + ky*3 + kx zero-padded from 36 to 48.  CUDA + fp16 only; elsewhere conv1 stays a convolution.
        self.gemm_conv1 = dev.type == "cuda" and dtype == torch.float16
        if self.gemm_conv1:
            w1p = torch.zeros((w1.shape[0], 48), dtype=torch.float32, device=dev)
            w1p[:, :36] = w1.reshape(w1.shape[0], 36)
            self.w1p_t = w1p.to(dtype).contiguous().t()               # [48,512] view of the row-major [512,48] weights
            from ... import _capi
            self._dll = _capi.lib()
            self._cols = {}
        self.convs = []
        for conv, bn, pad in ((net.conv2, net.bn2, 1), (net.conv3, net.bn3, 0), (net.conv4, net.bn4, 0)):
            w, b = _fold(conv.weight, conv.bias, bn)
            self.convs.append((w.to(dtype).contiguous(memory_format=torch.channels_last), b.to(dtype), pad))
        # fc1 consumes the NCHW flatten (c*9 + h*3 + w); permute its columns to the NHWC flatten we produce
        w, b = _fold(net.fc1.weight, net.fc1.bias, net.fc_bn1)
        ch = net.conv4.out_channels
        hw = net.flat // ch
        side = int(round(hw ** 0.5))
        w = w.view(-1, ch, side, side).permute(0, 2, 3, 1).reshape(-1, net.flat)
        self.fc1 = (w.to(dtype).contiguous(), b.to(dtype))
        w, b = _fold(net.fc2.weight, net.fc2.bias, net.fc_bn2)
        self.fc2 = (w.to(dtype).contiguous(), b.to(dtype))
        self.fc1_t, self.fc2_t = self.fc1[0].t(), self.fc2[0].t()
        # both heads in one GEMM: rows 0..342 = fc3, row 343 = fc4
        wh = torch.cat([net.fc3.weight.detach().float(), net.fc4.weight.detach().float()], dim=0)
        bh = torch.cat([net.fc3.bias.detach().float(), net.fc4.bias.detach().float()], dim=0)
        self.head = (wh.to(dtype).contiguous(), bh.to(dtype))
        self.head_t = self.head[0].t()
        self.n_act = net.action_size
        self.device = dev

    def _conv1_gemm(self, x: torch.Tensor):
        """conv1 + BN + ReLU as im2col (our kernel) + one GEMM with a fused bias+ReLU epilogue (cuBLASLt)."""
        n = x.shape[0]
        x = x.contiguous()
        # one im2col buffer per input buffer: two halves of a batch may be evaluated concurrently on two streams
        key = (x.data_ptr(), n)
        cols = self._cols.get(key)
        if cols is None:
            if len(self._cols) > 16:
                self._cols.clear()
            cols = self._cols[key] = torch.empty((n * 49, 48), dtype=self.dtype, device=x.device)
        _capi_check(self._dll.ifx_net_im2col3x3(x.data_ptr(), n, cols.data_ptr(),
                                                torch.cuda.current_stream(x.device).cuda_stream), self._dll)
        s = torch._addmm_activation(self.b1, cols, self.w1p_t, use_gelu=False)            # [n*49, 512] = NHWC
        return s.view(n, 7, 7, -1).permute(0, 3, 1, 2)                       # NCHW view of channels-last memory

    def _trunk(self, x: torch.Tensor):
        if self.gemm_conv1 and x.is_cuda and x.dtype == torch.float32 and tuple(x.shape[1:]) == (4, 7, 7):
            s = self._conv1_gemm(x)
        else:
            x = x.to(self.first_dtype).contiguous(memory_format=torch.channels_last)
            s = _conv_relu(x, self.w1, self.b1, 1).to(self.dtype)
        for w, b, pad in self.convs:
            s = _conv_relu(s, w, b, pad)
        s = s.permute(0, 2, 3, 1).reshape(s.shape[0], -1)                    # NHWC flatten: a view for channels-last
        if s.is_cuda and self.dtype != torch.float32:                        # bias + ReLU in the GEMM epilogue
            s = torch._addmm_activation(self.fc1[1], s, self.fc1_t, use_gelu=False)
            return torch._addmm_activation(self.fc2[1], s, self.fc2_t, use_gelu=False)
        s = F.relu(F.linear(s, *self.fc1))
        return F.relu(F.linear(s, *self.fc2))

    @torch.no_grad()
    def head_into(self, x: torch.Tensor, out: torch.Tensor):
        """Raw head output [B,344] (343 policy logits + value pre-activation) written into ``out`` by the last GEMM
        itself; softmax / tanh are left to the consumer (the tree kernel's fused-head mode)."""
        return torch.addmm(self.head[1], self._trunk(x), self.head_t, out=out)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor):
        out = F.linear(self._trunk(x), *self.head).float()
        pi = torch.exp(F.log_softmax(out[:, :self.n_act], dim=1))            # NNet.py:94
        v = torch.tanh(out[:, self.n_act])
        return pi, v
