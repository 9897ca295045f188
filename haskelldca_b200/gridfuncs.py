"""Gridfuncs operators (src/Gridfuncs.hs) on explicit arrays, executed on the GPU.

Same names and argument meaning as the reference; grids are uint8[7,7,70] (Bool),
freps int64[7,7,71] (Accelerate Int).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import CHANNELS, COLS, FREP_SIZE, GRID_SIZE, NFEATS, ROWS


def _grid(g):
    g = np.ascontiguousarray(g, dtype=np.uint8)
    if g.size != GRID_SIZE:
        raise ValueError("grid must have 7*7*70 elements")
    return g


def mk_grid():  # mkGrid, Gridfuncs.hs:27-30
    return np.zeros((ROWS, COLS, CHANNELS), dtype=np.uint8)


def _chs(fn, grid, cell, device):
    out = np.zeros(CHANNELS, dtype=np.int64)
    n = C.c_int32()
    _ffi.check(fn(device, _ffi.ptr(_grid(grid)), cell[0], cell[1], _ffi.ptr(out), C.byref(n)))
    return out[: n.value].copy()


def eligible_chs(cell, grid, device=0):  # Gridfuncs.hs:81-82
    return _chs(_ffi.load().dca_gf_eligible_chs, grid, cell, device)


def inuse_chs(cell, grid, device=0):  # Gridfuncs.hs:86-87
    return _chs(_ffi.load().dca_gf_inuse_chs, grid, cell, device)


def feature_rep(grid, device=0):  # Gridfuncs.hs:153-170
    out = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    _ffi.check(_ffi.load().dca_gf_feature_rep(device, _ffi.ptr(_grid(grid)), _ffi.ptr(out)))
    return out


def inc_afterstate_freps(cell, is_end, chs, grid, frep, device=0):  # Gridfuncs.hs:186-252
    chs = np.ascontiguousarray(chs, dtype=np.int64)
    frep = np.ascontiguousarray(frep, dtype=np.int64)
    if frep.size != FREP_SIZE:
        raise ValueError("frep must have 7*7*71 elements")
    out = np.zeros((len(chs), ROWS, COLS, NFEATS), dtype=np.int64)
    _ffi.check(_ffi.load().dca_gf_inc_afterstate_freps(
        device, cell[0], cell[1], int(is_end), _ffi.ptr(chs), len(chs), _ffi.ptr(_grid(grid)), _ffi.ptr(frep), _ffi.ptr(out)))
    return out


def violates_reuse_constraint(grid, device=0) -> bool:  # Gridfuncs.hs:94-101
    v = C.c_int32()
    _ffi.check(_ffi.load().dca_gf_violates_reuse_constraint(device, _ffi.ptr(_grid(grid)), C.byref(v)))
    return bool(v.value)
