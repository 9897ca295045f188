"""ctypes binding of libdca_b200.so (include/dca_b200.h).

The library is the product; this module only marshals numpy buffers into the C ABI.
If the shared library is missing the import fails loudly -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

ROWS, COLS, CHANNELS, NFEATS = 7, 7, 70, 71
GRID_SIZE, FREP_SIZE = 3430, 3479
ORDER_REF, ORDER_FAST = 0, 1
RNG_TAPE, RNG_PHILOX = 0, 1
NEW, END, HOFF = 0, 1, 2
RUNNING, SUCCESS, ZERO_LOSS, NAN_LOSS, REUSE_VIOLATED, NELIG_OVERFLOW, TAPE_EXHAUSTED, UNINITIALIZED, QUEUE_OVERFLOW = range(9)


class Config(C.Structure):
    _fields_ = [
        ("n_replicas", C.c_int32),
        ("device", C.c_int32),
        ("seed", C.c_uint64),
        ("first_replica", C.c_uint64),
        ("rng_mode", C.c_int32),
        ("order", C.c_int32),
        ("n_events", C.c_int64),
        ("min_loss", C.c_float),
        ("verify_reuse_constraint", C.c_int32),
        ("verify_nelig_bounds", C.c_int32),
        ("trace_capacity", C.c_int32),
        ("warps_per_replica", C.c_int32),
        ("call_dur", C.c_double),
        ("call_dur_hoff", C.c_double),
        ("call_rate", C.c_double),
        ("hoff_prob", C.c_double),
        ("alpha_net", C.c_float),
        ("alpha_avg", C.c_float),
        ("alpha_grad", C.c_float),
        ("call_dur_v", C.c_void_p),
        ("call_dur_hoff_v", C.c_void_p),
        ("call_rate_v", C.c_void_p),
        ("hoff_prob_v", C.c_void_p),
        ("alpha_net_v", C.c_void_p),
        ("alpha_avg_v", C.c_void_p),
        ("alpha_grad_v", C.c_void_p),
        ("hoff_lookahead", C.c_int32),
        ("reserved_", C.c_int32),
    ]


class Event(C.Structure):
    _fields_ = [
        ("time", C.c_double),
        ("etype", C.c_int32),
        ("r", C.c_int32),
        ("c", C.c_int32),
        ("ch", C.c_int32),
        ("hoff_r", C.c_int32),
        ("hoff_c", C.c_int32),
    ]

    def astuple(self):
        return (self.time, self.etype, self.r, self.c, self.ch, self.hoff_r, self.hoff_c)


TRACE_DTYPE = np.dtype(
    [
        ("time", "<f8"), ("loss", "<f4"), ("avg_reward", "<f4"), ("reward", "<i4"),
        ("etype", "u1"), ("cell", "u1"), ("ch", "i1"), ("ncand", "u1"), ("end_ch", "i1"),
        ("pad", "u1", (7,)), ("grid_hash", "<u8"), ("frep_hash", "<u8"),
    ]
)
assert TRACE_DTYPE.itemsize == 48

# every symbol include/dca_b200.h declares
SYMBOLS = {
    "dca_config_defaults": (C.c_int, [C.POINTER(Config)]),
    "dca_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "dca_destroy": (C.c_int, [C.c_void_p]),
    "dca_reset": (C.c_int, [C.c_void_p]),
    "dca_set_params": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7),
    "dca_last_error": (C.c_char_p, [C.c_void_p]),
    "dca_device_count": (C.c_int, []),
    "dca_set_rng_tape": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "dca_set_rng_mt19937": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_size_t]),
    "dca_run": (C.c_int, [C.c_void_p, C.c_int64]),
    "dca_run_async": (C.c_int, [C.c_void_p, C.c_int64]),
    "dca_wait": (C.c_int, [C.c_void_p]),
    "dca_run_many": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64]),
    "dca_run_unit_events": (C.c_int64, []),
    "dca_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "dca_last_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "dca_get_action": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "dca_grid_step_train": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "dca_acc_get_action": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_int32), C.c_void_p]),
    "dca_acc_grid_step_train": (C.c_int, [C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "dca_dense_dot": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "dca_checkpoint_size": (C.c_size_t, [C.c_void_p]),
    "dca_checkpoint_save": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dca_checkpoint_load": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dca_read_state": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(Event)]),
    "dca_write_state": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "dca_read_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "dca_read_status": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]),
    "dca_read_period": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.c_int32]),
    "dca_read_trace": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dca_read_summary": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dca_read_agents": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dca_sum_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "dca_stats_cumus": (None, [C.c_void_p, C.c_void_p]),
    "dca_get_neighs": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "dca_gf_eligible_chs": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "dca_gf_inuse_chs": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "dca_gf_feature_rep": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p]),
    "dca_gf_inc_afterstate_freps": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "dca_gf_violates_reuse_constraint": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "dca_allreduce_stats": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
}

_lib = None


def lib_path() -> str:
    """The product library; DCA_LIB selects a development variant built with build.build(out=...)."""
    return os.environ.get("DCA_LIB") or _build.LIB


def load(build_if_missing: bool = True):
    """Load libdca_b200.so; build it with nvcc if absent.  Never falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and path == _build.LIB and _build.is_stale():
        _build.build()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m haskelldca_b200.build` (nvcc, sm_100a)")
    L = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class DcaError(RuntimeError):
    pass


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().dca_last_error(handle)
        raise DcaError(f"libdca_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None
