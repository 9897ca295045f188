"""Replicas: a batch of independent simulator states on one GPU (SimState, src/Simulator.hs:59-97)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import CHANNELS, COLS, FREP_SIZE, NFEATS, ROWS
from .opt import Opt


class Replicas:
    """mkSimState for n replicas + the operators that act on them.

    Simulation parameters may be given per replica (arrays of n_replicas values)
    through `per_replica`, keyed by the Opt field names call_dur, call_dur_hoff,
    call_rate, hoff_prob, learning_rate, learning_rate_avg, learning_rate_grad.
    """

    def __init__(self, opt: Opt | None = None, per_replica: dict | None = None, **kw):
        opt = opt or Opt()
        for k, v in kw.items():
            if not hasattr(opt, k):
                raise AttributeError(k)
            setattr(opt, k, v)
        self.opt = opt
        L = _ffi.load()
        cfg = _ffi.Config()
        _ffi.check(L.dca_config_defaults(C.byref(cfg)))
        cfg.n_replicas = opt.n_replicas
        cfg.device = opt.device
        cfg.seed = opt.rng_seed
        cfg.first_replica = opt.first_replica
        cfg.rng_mode = _ffi.RNG_PHILOX if opt.rng == "philox" else _ffi.RNG_TAPE
        cfg.order = _ffi.ORDER_FAST if opt.order == "fast" else _ffi.ORDER_REF
        cfg.n_events = opt.n_events
        cfg.min_loss = opt.min_loss
        cfg.verify_reuse_constraint = int(opt.verify_reuse_constraint)
        cfg.verify_nelig_bounds = int(opt.verify_nelig_bounds)
        cfg.trace_capacity = opt.trace
        cfg.warps_per_replica = opt.warps_per_replica
        cfg.hoff_lookahead = int(opt.hoff_lookahead)
        cfg.call_dur, cfg.call_dur_hoff = opt.call_dur, opt.call_dur_hoff
        cfg.call_rate, cfg.hoff_prob = opt.call_rate, opt.hoff_prob
        cfg.alpha_net, cfg.alpha_avg, cfg.alpha_grad = opt.learning_rate, opt.learning_rate_avg, opt.learning_rate_grad
        self._keep = []
        names = {"call_dur": ("call_dur_v", np.float64), "call_dur_hoff": ("call_dur_hoff_v", np.float64),
                 "call_rate": ("call_rate_v", np.float64), "hoff_prob": ("hoff_prob_v", np.float64),
                 "learning_rate": ("alpha_net_v", np.float32), "learning_rate_avg": ("alpha_avg_v", np.float32),
                 "learning_rate_grad": ("alpha_grad_v", np.float32)}
        self.h2d_bytes = 0
        for k, v in (per_replica or {}).items():
            field, dt = names[k]
            arr = np.ascontiguousarray(v, dtype=dt)
            if arr.shape != (opt.n_replicas,):
                raise ValueError(f"per_replica[{k}] must have n_replicas entries")
            self._keep.append(arr)
            self.h2d_bytes += arr.nbytes
            setattr(cfg, field, arr.ctypes.data)
        self.n = opt.n_replicas
        self.h = C.c_void_p()
        _ffi.check(L.dca_create(C.byref(cfg), C.byref(self.h)))
        if opt.rng == "mt19937":  # --fixed_rng replay: MT19937-64 stream of the reference (EventGen/Internal.hs:57-66)
            n_draws = 4 * opt.n_events + 64
            for r in range(self.n):
                self.set_rng_mt19937(r, opt.rng_seed + r, n_draws)

    # ---- lifecycle -----------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            _ffi.load().dca_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self):
        _ffi.check(_ffi.load().dca_reset(self.h), self.h)

    def set_params(self, **per_replica):
        """Upload new per-replica parameter arrays (effective at the next reset); returns bytes copied."""
        order = [("call_dur", np.float64), ("call_dur_hoff", np.float64), ("call_rate", np.float64), ("hoff_prob", np.float64),
                 ("learning_rate", np.float32), ("learning_rate_avg", np.float32), ("learning_rate_grad", np.float32)]
        args, nbytes = [], 0
        for name, dt in order:
            v = per_replica.get(name)
            if v is None:
                args.append(None)
                continue
            arr = np.ascontiguousarray(v, dtype=dt)
            if arr.shape != (self.n,):
                raise ValueError(f"{name} must have n_replicas entries")
            args.append(arr)
            nbytes += arr.nbytes
        _ffi.check(_ffi.load().dca_set_params(self.h, *[_ffi.ptr(a) for a in args]), self.h)
        return nbytes

    def set_rng_tape(self, replica, words, neglog):
        words = np.ascontiguousarray(words, dtype=np.uint64)
        neglog = np.ascontiguousarray(neglog, dtype=np.float64)
        _ffi.check(_ffi.load().dca_set_rng_tape(self.h, replica, _ffi.ptr(words), _ffi.ptr(neglog), len(words)), self.h)

    def set_rng_mt19937(self, replica, seed, n_draws):
        _ffi.check(_ffi.load().dca_set_rng_mt19937(self.h, replica, seed, n_draws), self.h)

    # ---- fused run (runPeriod, SimRunner.hs:132-178) ---------------------------
    def run(self, max_events: int) -> float:
        """Advance every running replica by up to max_events events; returns device ms."""
        _ffi.check(_ffi.load().dca_run(self.h, int(max_events)), self.h)
        return self.elapsed_ms()

    def run_async(self, max_events: int) -> None:
        """Enqueue the launch and return; `wait` (or any other call on this object) collects it."""
        _ffi.check(_ffi.load().dca_run_async(self.h, int(max_events)), self.h)

    def wait(self) -> float:
        _ffi.check(_ffi.load().dca_wait(self.h), self.h)
        return self.elapsed_ms()

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        _ffi.check(_ffi.load().dca_elapsed_ms(self.h, C.byref(ms)), self.h)
        return ms.value

    def launch_count(self) -> int:
        n = C.c_int64()
        _ffi.check(_ffi.load().dca_last_launch_count(self.h, C.byref(n)), self.h)
        return n.value

    # ---- step-wise operators (accFn1 / accFn2, SimRunner.hs:61-62) -------------
    def get_action(self, replica, cell, is_end) -> int:
        ch = C.c_int32()
        _ffi.check(_ffi.load().dca_get_action(self.h, replica, cell[0], cell[1], int(is_end), C.byref(ch)), self.h)
        return ch.value

    def grid_step_train(self, replica, is_end, cell, ch, alpha_net=None, alpha_avg=None, alpha_grad=None):
        o = self.opt
        loss, isnan, reward = C.c_float(), C.c_int32(), C.c_int64()
        _ffi.check(_ffi.load().dca_grid_step_train(
            self.h, replica,
            o.learning_rate if alpha_net is None else alpha_net,
            o.learning_rate_avg if alpha_avg is None else alpha_avg,
            o.learning_rate_grad if alpha_grad is None else alpha_grad,
            int(is_end), cell[0], cell[1], int(ch), C.byref(loss), C.byref(isnan), C.byref(reward)), self.h)
        return np.float32(loss.value), bool(isnan.value), reward.value

    # ---- checkpoint / resume ----------------------------------------------------------
    def checkpoint(self) -> np.ndarray:
        """the complete state of every replica as one opaque uint8 blob"""
        L = _ffi.load()
        buf = np.empty(L.dca_checkpoint_size(self.h), dtype=np.uint8)
        _ffi.check(L.dca_checkpoint_save(self.h, _ffi.ptr(buf), buf.nbytes), self.h)
        return buf

    def restore(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        _ffi.check(_ffi.load().dca_checkpoint_load(self.h, _ffi.ptr(blob), blob.nbytes), self.h)

    # ---- read-back ---------------------------------------------------------------
    def read_state(self, replica=0):
        grid = np.zeros((ROWS, COLS, CHANNELS), dtype=np.uint8)
        frep = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
        w = np.zeros(FREP_SIZE, dtype=np.float32)
        wg = np.zeros(FREP_SIZE, dtype=np.float32)
        avg, it, ev = C.c_float(), C.c_int64(), _ffi.Event()
        _ffi.check(_ffi.load().dca_read_state(self.h, replica, _ffi.ptr(grid), _ffi.ptr(frep), _ffi.ptr(w), _ffi.ptr(wg),
                                              C.byref(avg), C.byref(it), C.byref(ev)), self.h)
        return dict(grid=grid, frep=frep, w=w, wg=wg, avg_reward=np.float32(avg.value), iter=it.value,
                    event=ev.astuple(), stats=self.read_stats(replica))

    def write_state(self, replica, grid=None, w=None, wg=None, avg_reward=None):
        g = np.ascontiguousarray(grid, dtype=np.uint8) if grid is not None else None
        w = np.ascontiguousarray(w, dtype=np.float32) if w is not None else None
        wg = np.ascontiguousarray(wg, dtype=np.float32) if wg is not None else None
        avg = C.byref(C.c_float(avg_reward)) if avg_reward is not None else None
        _ffi.check(_ffi.load().dca_write_state(self.h, replica, _ffi.ptr(g), _ffi.ptr(w), _ffi.ptr(wg), avg), self.h)

    def read_stats(self, replica=0):
        st = np.zeros(8, dtype=np.int64)
        _ffi.check(_ffi.load().dca_read_stats(self.h, replica, _ffi.ptr(st)), self.h)
        return st

    def read_status(self, replica=0):
        s, it, nd = C.c_int32(), C.c_int64(), C.c_uint64()
        _ffi.check(_ffi.load().dca_read_status(self.h, replica, C.byref(s), C.byref(it), C.byref(nd)), self.h)
        return s.value, it.value, nd.value

    def read_period(self, replica=0, reset=True):
        s, n = C.c_float(), C.c_int64()
        _ffi.check(_ffi.load().dca_read_period(self.h, replica, C.byref(s), C.byref(n), int(reset)), self.h)
        return np.float32(s.value), n.value

    def read_trace(self, replica=0):
        cap = self.opt.trace
        out = np.zeros(cap, dtype=_ffi.TRACE_DTYPE)
        n = C.c_size_t()
        _ffi.check(_ffi.load().dca_read_trace(self.h, replica, _ffi.ptr(out), cap, C.byref(n)), self.h)
        return out[: n.value]

    def read_summary(self):
        status = np.zeros(self.n, dtype=np.int32)
        it = np.zeros(self.n, dtype=np.int64)
        avg = np.zeros(self.n, dtype=np.float32)
        st = np.zeros((self.n, 8), dtype=np.int64)
        _ffi.check(_ffi.load().dca_read_summary(self.h, _ffi.ptr(status), _ffi.ptr(it), _ffi.ptr(avg), _ffi.ptr(st)), self.h)
        return dict(status=status, iter=it, avg_reward=avg, stats=st)

    def read_agents(self, out=None):
        """The learned agents of all replicas: (wnet [n, 3479], wgrad [n, 3479], avg_reward [n]) in the reference's
        flatten order.  `out` = preallocated (pinned) arrays to fill, e.g. from a previous call."""
        if out is None:
            out = (np.empty((self.n, FREP_SIZE), np.float32), np.empty((self.n, FREP_SIZE), np.float32), np.empty(self.n, np.float32))
        w, wg, avg = out
        _ffi.check(_ffi.load().dca_read_agents(self.h, _ffi.ptr(w), _ffi.ptr(wg), _ffi.ptr(avg)), self.h)
        return w, wg, avg

    def sum_stats(self):
        """int64[10]: summed Stats counters, events processed, failed replicas; plus its device address."""
        out = np.zeros(10, dtype=np.int64)
        dev = C.c_void_p()
        _ffi.check(_ffi.load().dca_sum_stats(self.h, _ffi.ptr(out), C.byref(dev)), self.h)
        return out, dev.value


def run_many(replica_sets, max_events: int) -> list:
    """dca_run_many: one launch per `Replicas` (typically one per device of the box), all in flight together although
    the caller is a single thread; returns the device ms of each."""
    n = len(replica_sets)
    arr = (C.c_void_p * n)(*[s.h for s in replica_sets])
    _ffi.check(_ffi.load().dca_run_many(arr, n, int(max_events)), replica_sets[0].h)
    return [s.elapsed_ms() for s in replica_sets]


def allreduce_stats(replica_sets) -> np.ndarray:
    """dca_allreduce_stats: the int64[10] of `sum_stats` summed over several `Replicas` of THIS process, one per
    device, with ncclAllReduce over NVLink (the only collective of the path: the blocking-probability counters,
    src/Stats.hs:83-92).  With one process per GPU all-reduce `sum_stats()[1]` with the process group instead."""
    L = _ffi.load()
    n = len(replica_sets)
    arr = (C.c_void_p * n)(*[s.h for s in replica_sets])
    out = np.zeros(10, dtype=np.int64)
    _ffi.check(L.dca_allreduce_stats(arr, n, _ffi.ptr(out)), replica_sets[0].h)
    return out
