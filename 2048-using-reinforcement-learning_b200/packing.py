"""Board formats: the reference's int32[16] tile values  <->  packed uint64 nibble exponents.

Cell (r, c) is flat index i = 4r + c (row-major, `board.flatten()` of environment/game_2048.py:57);
nibble i of the packed word holds log2(tile), 0 for an empty cell.
"""
from __future__ import annotations

import numpy as np

_SHIFTS = (4 * np.arange(16, dtype=np.uint64)).reshape(1, 16)


def pack_boards(values) -> np.ndarray:
    """int[..., 16] (or [..., 4, 4]) tile values -> uint64[...] packed boards."""
    v = np.asarray(values)
    if v.shape[-2:] == (4, 4):
        v = v.reshape(v.shape[:-2] + (16,))
    if v.shape[-1] != 16:
        raise ValueError(f"expected trailing dimension 16 or (4, 4), got {v.shape}")
    v64 = v.astype(np.int64, copy=False)
    if (v64 < 0).any() or (v64 & (v64 - 1)).any() or (v64 == 1).any():
        raise ValueError("tile values must be 0 or powers of two >= 2")
    if (v64 > 32768).any():
        raise ValueError("tiles above 32768 do not fit the 4-bit exponent format")
    e = np.zeros(v64.shape, dtype=np.uint64)
    nz = v64 > 0
    e[nz] = np.log2(v64[nz]).astype(np.uint64)
    flat = e.reshape(-1, 16)
    packed = (flat << _SHIFTS).sum(axis=1, dtype=np.uint64)
    return packed.reshape(v.shape[:-1])


def unpack_boards(packed) -> np.ndarray:
    """uint64[...] packed boards -> int32[..., 16] tile values."""
    p = np.asarray(packed, dtype=np.uint64)
    e = (p.reshape(-1, 1) >> _SHIFTS) & np.uint64(15)
    vals = np.where(e > 0, np.left_shift(np.int64(1), e.astype(np.int64)), 0).astype(np.int32)
    return vals.reshape(p.shape + (16,))


def pack_board(values) -> int:
    return int(pack_boards(np.asarray(values).reshape(1, -1))[0])


def unpack_board(packed: int) -> np.ndarray:
    return unpack_boards(np.array([packed], dtype=np.uint64))[0]
