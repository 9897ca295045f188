"""Gridneighs operators (src/Gridneighs.hs:92-98) through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def get_neighs(d: int, cell, include_self: bool):
    """getNeighs d cell includeself -> list of (r, c) in the reference's ring order."""
    buf = np.zeros(2 * 64, dtype=np.int32)
    n = C.c_int32()
    _ffi.check(_ffi.load().dca_get_neighs(d, cell[0], cell[1], int(include_self), _ffi.ptr(buf), C.byref(n)))
    return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)]
