"""Packed 64-byte game state (``ifx_state`` in include/ifx_b200.h) <-> the reference's board arrays.

Host-side conversion only (numpy bit operations); the rules themselves run on the device.
Reference attributes replaced: ``InflexionGame._board/_player/_curr_turn/_outcome/_max_turns``
(inflexion/InflexionGame.py:57-69, Game.py:21-25).
"""
from __future__ import annotations

import numpy as np

STATE_DTYPE = np.dtype([
    ("red", "<u8"), ("blue", "<u8"), ("p0", "<u8"), ("p1", "<u8"), ("p2", "<u8"),
    ("turn", "<u2"), ("max_turns", "<u2"), ("player", "u1"), ("outcome", "u1"),
    ("total_power", "u1"), ("flags", "u1"), ("tag", "<u4"), ("reserved", "<u4", (3,)),
])
assert STATE_DTYPE.itemsize == 64

N, CELLS, MOVES, ACTIONS = 7, 49, 7, 343
ONGOING, DRAW, WON, LOST = 0, 1, 2, 3
_BITS = (np.uint64(1) << np.arange(CELLS, dtype=np.uint64))


def _to_mask(bits: np.ndarray) -> np.ndarray:
    """bool [S,49] -> uint64 [S]"""
    return (bits.astype(np.uint64) * _BITS).sum(axis=1, dtype=np.uint64)


def _from_mask(mask: np.ndarray) -> np.ndarray:
    """uint64 [S] -> uint8 [S,49]"""
    return ((mask[:, None] >> np.arange(CELLS, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)


def pack_states(boards, players, turns, outcomes=0, max_turns=343, tags=0) -> np.ndarray:
    """boards [S,49] or [S,7,7] signed powers; players +1 (RED) / -1 (BLUE); -> STATE_DTYPE [S]."""
    b = np.asarray(boards).reshape(-1, CELLS).astype(np.int64)
    s = len(b)
    out = np.zeros(s, dtype=STATE_DTYPE)
    mag = np.abs(b)
    if mag.max(initial=0) > 6:
        raise ValueError("stack power out of range [-6, 6]")
    out["red"] = _to_mask(b > 0)
    out["blue"] = _to_mask(b < 0)
    out["p0"] = _to_mask((mag & 1) > 0)
    out["p1"] = _to_mask((mag & 2) > 0)
    out["p2"] = _to_mask((mag & 4) > 0)
    out["turn"] = np.broadcast_to(np.asarray(turns), (s,))
    out["max_turns"] = np.broadcast_to(np.asarray(max_turns), (s,))
    out["player"] = (np.broadcast_to(np.asarray(players), (s,)) < 0).astype(np.uint8)
    out["outcome"] = np.broadcast_to(np.asarray(outcomes), (s,))
    out["total_power"] = mag.sum(axis=1)
    out["tag"] = np.broadcast_to(np.asarray(tags), (s,))
    return out


def unpack_boards(states: np.ndarray) -> np.ndarray:
    """STATE_DTYPE [S] -> int64 [S,49] signed powers."""
    mag = (_from_mask(states["p0"]).astype(np.int64) + 2 * _from_mask(states["p1"]) + 4 * _from_mask(states["p2"]))
    sign = _from_mask(states["red"]).astype(np.int64) - _from_mask(states["blue"]).astype(np.int64)
    return mag * sign


def players_of(states: np.ndarray) -> np.ndarray:
    return np.where(states["player"] == 0, 1, -1).astype(np.int64)


def unpack_masks(masks: np.ndarray) -> np.ndarray:
    """uint64 [S,7] per-move-type cell masks -> uint8 [S,343] (valid_actions_mask layout, a = m*49 + r*7 + q)."""
    m = np.asarray(masks, dtype=np.uint64).reshape(-1, MOVES)
    bits = (m[:, :, None] >> np.arange(CELLS, dtype=np.uint64)) & np.uint64(1)
    return bits.reshape(len(m), ACTIONS).astype(np.uint8)
