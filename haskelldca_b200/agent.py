"""Agent / Simulator operators on explicit arrays (src/Agent.hs, src/Simulator.hs:137-205,
src/SimRunner.hs:100-118), executed on the GPU.

These are the pure array functions the reference compiles with `runN`:
`get_action` is accFn1, `grid_step_train` is accFn2 (src/SimRunner.hs:61-62, 211-212).
Layouts as in the reference: grid uint8[7,7,70], frep int64[7,7,71], weights float32[3479].
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import COLS, FREP_SIZE, GRID_SIZE, NFEATS, ROWS

ORDERS = {"ref": _ffi.ORDER_REF, "fast": _ffi.ORDER_FAST}


def _order(order):
    return ORDERS[order] if isinstance(order, str) else int(order)


def _frep(f):
    f = np.ascontiguousarray(f, dtype=np.int64)
    if f.size != FREP_SIZE:
        raise ValueError("frep must have 7*7*71 elements")
    return f


def _w(w):
    w = np.ascontiguousarray(w, dtype=np.float32)
    if w.size != FREP_SIZE:
        raise ValueError("weight vectors have 7*7*71 elements")
    return w


def _grid(g):
    g = np.ascontiguousarray(g, dtype=np.uint8)
    if g.size != GRID_SIZE:
        raise ValueError("grid must have 7*7*70 elements")
    return g


def forward(frep, w_net, order="ref", device=0) -> np.float32:  # Agent.hs:22-23
    out = C.c_float()
    _ffi.check(_ffi.load().dca_dense_dot(device, _order(order), 0, _ffi.ptr(_frep(frep)), _ffi.ptr(_w(w_net)), C.byref(out)))
    return np.float32(out.value)


def dot_grad_corr(frep, w_grad, order="ref", device=0) -> np.float32:  # vvMul x wGradCorr, Agent.hs:66
    out = C.c_float()
    _ffi.check(_ffi.load().dca_dense_dot(device, _order(order), 1, _ffi.ptr(_frep(frep)), _ffi.ptr(_w(w_grad)), C.byref(out)))
    return np.float32(out.value)


def get_action(cell, is_end, w_net, grid, frep, order="ref", device=0):
    """accFn1 = getAction (Simulator.hs:154-172): -> (ch or -1, next_frep)."""
    nf = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    ch = C.c_int32()
    _ffi.check(_ffi.load().dca_acc_get_action(device, _order(order), cell[0], cell[1], int(is_end), _ffi.ptr(_w(w_net)),
                                              _ffi.ptr(_grid(grid)), _ffi.ptr(_frep(frep)), C.byref(ch), _ffi.ptr(nf)))
    return ch.value, nf


def grid_step_train(alpha_net, alpha_avg, alpha_grad, is_end, cell, ch, frep, next_frep, grid, agent, order="ref", device=0):
    """accFn2 = gridStepTrain (SimRunner.hs:100-118); agent = (w_net, w_grad_corr, avg_reward).
    -> (next_grid, next_agent, loss or None when NaN, reward)."""
    w_net, w_grad, avg = agent
    g = _grid(grid).copy()
    w = _w(w_net).copy()
    wg = _w(w_grad).copy()
    avg_c, loss, isnan, reward = C.c_float(float(avg)), C.c_float(), C.c_int32(), C.c_int64()
    _ffi.check(_ffi.load().dca_acc_grid_step_train(
        device, _order(order), alpha_net, alpha_avg, alpha_grad, int(is_end), cell[0], cell[1], -1 if ch is None else int(ch),
        _ffi.ptr(_frep(frep)), _ffi.ptr(_frep(next_frep)), _ffi.ptr(g), _ffi.ptr(w), _ffi.ptr(wg),
        C.byref(avg_c), C.byref(loss), C.byref(isnan), C.byref(reward)))
    return g.reshape(np.shape(grid)), (w, wg, np.float32(avg_c.value)), (None if isnan.value else np.float32(loss.value)), reward.value
