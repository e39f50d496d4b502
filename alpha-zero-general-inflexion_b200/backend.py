"""Where the kernels run: the CUDA library on a CUDA device.  There is no other product backend.

``Backend.cuda()`` is what every public class uses by default; it raises when the library or the device is
missing.  The constructor also accepts an explicitly injected (dll, device) pair -- a test seam used by
tests/ to drive the GPU-less logic-check build of the same kernels; nothing in the package selects it.
"""
from __future__ import annotations

import torch

from . import _capi


class Backend:
    def __init__(self, dll, device):
        self.dll = dll
        self.device = torch.device(device)
        self.is_cuda = self.device.type == "cuda"

    @staticmethod
    def cuda(index: int | None = None) -> "Backend":
        dll = _capi.lib()                                   # raises if libifx_b200.so is not built
        if not torch.cuda.is_available():
            raise _capi.IfxError("inflexion_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if index is None:
            index = torch.cuda.current_device()
        return Backend(dll, torch.device("cuda", index))

    def stream(self):
        """cudaStream_t of torch's current stream (so launches order with torch ops and can be graph-captured)."""
        if self.is_cuda:
            return torch.cuda.current_stream(self.device).cuda_stream
        return None

    def check(self, rc: int) -> None:
        _capi.check(rc, self.dll)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def states_tensor(self, states) -> torch.Tensor:
        """numpy STATE_DTYPE [S] or uint8 tensor [S,64] -> uint8 tensor [S,64] on the device."""
        import numpy as np
        if isinstance(states, torch.Tensor):
            t = states
        else:
            t = torch.from_numpy(np.ascontiguousarray(states).view(np.uint8).reshape(-1, 64))
        return t.to(self.device).contiguous()


_DEFAULT = None


def default_backend() -> Backend:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Backend.cuda()
    return _DEFAULT
