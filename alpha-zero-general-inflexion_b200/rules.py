"""Batched Inflexion rules on the device: thin tensor wrappers over ifx_rules_* (include/ifx_b200.h).

States are uint8 tensors [S,64] holding packed ``ifx_state`` records (see state.py for the numpy view).
Replaces, for S states at once: ``valid_actions_mask`` / ``to_next_state`` / ``to_planes`` /
``random_symmetry`` / ``symmetries`` / ``score`` of inflexion/InflexionGame.py:75-122,198.
"""
from __future__ import annotations

import numpy as np
import torch

from .backend import Backend, default_backend
from .state import STATE_DTYPE


def _ptr(t):
    return None if t is None else t.data_ptr()


def mask_step(states: torch.Tensor, actions: torch.Tensor | None, want_masks: bool = True,
              backend: Backend | None = None):
    """-> (masks int64 [S,7] or None, next_states uint8 [S,64] or None, status int32 [S] or None)."""
    be = backend or default_backend()
    s = states.shape[0]
    assert states.dtype == torch.uint8 and states.shape[1:] == (64,) and states.is_contiguous()
    masks = be.empty((s, 7), torch.int64) if want_masks else None
    nxt = status = None
    if actions is not None:
        assert actions.dtype == torch.int32 and actions.shape == (s,) and actions.is_contiguous()
        nxt = be.empty((s, 64), torch.uint8)
        status = be.empty((s,), torch.int32)
    be.check(be.dll.ifx_rules_mask_step(_ptr(states), _ptr(actions), s, _ptr(masks), _ptr(nxt), _ptr(status),
                                        be.stream()))
    return masks, nxt, status


def planes(states: torch.Tensor, sym: torch.Tensor | None = None, dtype=torch.float32,
           backend: Backend | None = None) -> torch.Tensor:
    """to_planes for every state -> [S,4,7,7]; sym int32 [S] = k + 6*shift + 42*axis applies random_symmetry."""
    be = backend or default_backend()
    s = states.shape[0]
    assert dtype in (torch.float32, torch.int64)
    out = be.empty((s, 4, 7, 7), dtype)
    be.check(be.dll.ifx_rules_planes(_ptr(states), _ptr(sym), s, 0 if dtype == torch.float32 else 1, _ptr(out),
                                     be.stream()))
    return out


def symmetries(x: torch.Tensor, backend: Backend | None = None) -> torch.Tensor:
    """[S,C,7,7] (4- or 8-byte elements) -> [S,36,C,7,7] in the order of InflexionGame.symmetries."""
    be = backend or default_backend()
    assert x.dim() == 4 and x.shape[2:] == (7, 7) and x.is_contiguous() and x.element_size() in (4, 8)
    s, c = x.shape[:2]
    out = be.empty((s, 36, c, 7, 7), x.dtype)
    be.check(be.dll.ifx_rules_symmetries(_ptr(x), s, c, x.element_size(), _ptr(out), be.stream()))
    return out


def transform(x: torch.Tensor, sym: torch.Tensor, backend: Backend | None = None) -> torch.Tensor:
    """[S,C,7,7] -> translate(rotate(x, k), shift, axis) per item, sym int32 [S] = k + 6*shift + 42*axis."""
    be = backend or default_backend()
    assert x.dim() == 4 and x.shape[2:] == (7, 7) and x.is_contiguous() and x.element_size() in (4, 8)
    out = torch.empty_like(x)
    be.check(be.dll.ifx_rules_transform(_ptr(x), _ptr(sym), x.shape[0], x.shape[1], x.element_size(), _ptr(out),
                                        be.stream()))
    return out


def random_actions(states: torch.Tensor, seed: int, step: int, backend: Backend | None = None) -> torch.Tensor:
    """RandomPlayer for every state: uniform over the valid actions (stream = state tag) -> int32 [S], -1 = game over."""
    be = backend or default_backend()
    out = be.empty((states.shape[0],), torch.int32)
    be.check(be.dll.ifx_rules_random_actions(_ptr(states), states.shape[0], int(seed) & (2 ** 64 - 1), int(step),
                                             _ptr(out), be.stream()))
    return out


def greedy_actions(states: torch.Tensor, backend: Backend | None = None) -> torch.Tensor:
    """GreedyPlayer for every state: best one-ply score(), ties to the highest action -> int32 [S], -1 = game over."""
    be = backend or default_backend()
    out = be.empty((states.shape[0],), torch.int32)
    be.check(be.dll.ifx_rules_greedy_actions(_ptr(states), states.shape[0], _ptr(out), be.stream()))
    return out


def info(states: torch.Tensor, backend: Backend | None = None) -> torch.Tensor:
    """-> int32 [S,4]: score, power_diff (side to move), total_power, number of valid actions."""
    be = backend or default_backend()
    out = be.empty((states.shape[0], 4), torch.int32)
    be.check(be.dll.ifx_rules_info(_ptr(states), states.shape[0], _ptr(out), be.stream()))
    return out


# ---- host-buffer forms (numpy in, numpy out; the copies happen inside the library) --------------------

def _np_ptr(a):
    return None if a is None else a.ctypes.data


def mask_step_host(states: np.ndarray, actions: np.ndarray | None, want_masks: bool = True, backend=None):
    be = backend or default_backend()
    states = np.ascontiguousarray(states, dtype=STATE_DTYPE)
    s = len(states)
    masks = np.empty((s, 7), dtype=np.uint64) if want_masks else None
    nxt = status = None
    if actions is not None:
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        nxt = np.empty(s, dtype=STATE_DTYPE)
        status = np.empty(s, dtype=np.int32)
    be.check(be.dll.ifx_rules_mask_step_host(_np_ptr(states), _np_ptr(actions), s, _np_ptr(masks), _np_ptr(nxt),
                                             _np_ptr(status)))
    return masks, nxt, status


def planes_host(states: np.ndarray, sym: np.ndarray | None = None, dtype=np.int64, backend=None) -> np.ndarray:
    be = backend or default_backend()
    states = np.ascontiguousarray(states, dtype=STATE_DTYPE)
    if sym is not None:
        sym = np.ascontiguousarray(sym, dtype=np.int32)
    out = np.empty((len(states), 4, 7, 7), dtype=dtype)
    assert out.dtype in (np.float32, np.int64)
    be.check(be.dll.ifx_rules_planes_host(_np_ptr(states), _np_ptr(sym), len(states),
                                          0 if out.dtype == np.float32 else 1, _np_ptr(out)))
    return out


def symmetries_host(x: np.ndarray, backend=None) -> np.ndarray:
    be = backend or default_backend()
    x = np.ascontiguousarray(x)
    assert x.ndim == 4 and x.shape[2:] == (7, 7) and x.itemsize in (4, 8)
    out = np.empty((x.shape[0], 36) + x.shape[1:], dtype=x.dtype)
    be.check(be.dll.ifx_rules_symmetries_host(_np_ptr(x), x.shape[0], x.shape[1], x.itemsize, _np_ptr(out)))
    return out


def transform_host(x: np.ndarray, sym, backend=None) -> np.ndarray:
    be = backend or default_backend()
    x = np.ascontiguousarray(x)
    assert x.ndim == 4 and x.shape[2:] == (7, 7) and x.itemsize in (4, 8)
    sym = np.ascontiguousarray(sym, dtype=np.int32).reshape(-1)
    assert len(sym) == x.shape[0]
    out = np.empty_like(x)
    be.check(be.dll.ifx_rules_transform_host(_np_ptr(x), _np_ptr(sym), x.shape[0], x.shape[1], x.itemsize, _np_ptr(out)))
    return out


def info_host(states: np.ndarray, backend=None) -> np.ndarray:
    be = backend or default_backend()
    states = np.ascontiguousarray(states, dtype=STATE_DTYPE)
    out = np.empty((len(states), 4), dtype=np.int32)
    be.check(be.dll.ifx_rules_info_host(_np_ptr(states), len(states), _np_ptr(out)))
    return out
