"""ctypes binding of the CPU ORACLE (oracle/dca_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(haskelldca_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdca_oracle.so")

ROWS, COLS, CHANNELS, NFEATS = 7, 7, 70, 71
GRID_SIZE, FREP_SIZE = 3430, 3479
ORDER_REF, ORDER_FAST = 0, 1
RNG_MT, RNG_PHILOX = 0, 1
NEW, END, HOFF = 0, 1, 2
RUNNING, SUCCESS, ZERO_LOSS, NAN_LOSS, REUSE_VIOLATED, NELIG_OVERFLOW, HOST_ERROR = range(7)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, no reference sources)."""
    srcs = [os.path.join(_HERE, f) for f in ("dca_oracle.c", "dca_cpu_fast.c", "dca_oracle.h", "Makefile")]
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs)
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Opt(C.Structure):
    _fields_ = [
        ("call_dur", C.c_double),
        ("call_dur_hoff", C.c_double),
        ("call_rate", C.c_double),
        ("hoff_prob", C.c_double),
        ("n_events", C.c_int64),
        ("log_iter", C.c_int64),
        ("alpha_net", C.c_float),
        ("alpha_avg", C.c_float),
        ("alpha_grad", C.c_float),
        ("min_loss", C.c_float),
        ("verify_reuse_constraint", C.c_int32),
        ("verify_nelig_bounds", C.c_int32),
        ("seed", C.c_uint64),
        ("order", C.c_int32),
        ("rng_mode", C.c_int32),
        ("replica", C.c_uint64),
        ("hoff_lookahead", C.c_int32),
        ("pad_", C.c_int32),
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
        ("time", "<f8"),
        ("loss", "<f4"),
        ("avg_reward", "<f4"),
        ("reward", "<i4"),
        ("etype", "u1"),
        ("cell", "u1"),
        ("ch", "i1"),
        ("ncand", "u1"),
        ("end_ch", "i1"),
        ("pad", "u1", (7,)),
        ("grid_hash", "<u8"),
        ("frep_hash", "<u8"),
    ]
)
assert TRACE_DTYPE.itemsize == 48

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, f32, f64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64
        sig = {
            "orc_opt_defaults": (None, [C.POINTER(Opt)]),
            "orc_periphery": (i32, [i32, i32, i32, vp]),
            "orc_get_neighs": (i32, [i32, i32, i32, i32, vp]),
            "orc_get_neighborhood_acc": (i32, [i32, i32, i32, i32, vp]),
            "orc_table_sizes": (None, [vp, vp, vp, vp]),
            "orc_eligible_chs": (i32, [vp, i32, i32, vp]),
            "orc_inuse_chs": (i32, [vp, i32, i32, vp]),
            "orc_violates_reuse_constraint": (i32, [vp]),
            "orc_execute_action": (None, [i32, i32, i32, i32, vp]),
            "orc_afterstates": (None, [vp, i32, i32, i32, vp, i32, vp]),
            "orc_feature_rep": (None, [vp, vp]),
            "orc_feature_rep_slow": (None, [vp, vp]),
            "orc_inc_afterstate_freps": (None, [i32, i32, i32, vp, i32, vp, vp, vp]),
            "orc_bool_sum": (i64, [vp]),
            "orc_dense_dot": (f32, [i32, vp, vp]),
            "orc_dense_dot_g": (f32, [i32, vp, vp]),
            "orc_forward": (f32, [i32, vp, vp]),
            "orc_argpmax": (i32, [vp, i32]),
            "orc_select_action": (i32, [i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp]),
            "orc_get_action": (i32, [i32, i32, i32, i32, vp, vp, vp, vp, vp]),
            "orc_get_action_lookahead": (i32, [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
            "orc_backward": (f32, [i32, f32, f32, f32, vp, vp, f32, vp, vp, vp, vp]),
            "orc_grid_step_train": (f32, [i32, f32, f32, f32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
            "orc_rng_create": (vp, [i32, u64, u64]),
            "orc_rng_destroy": (None, [vp]),
            "orc_rng_word64": (u64, [vp]),
            "orc_rng_double": (f64, [vp]),
            "orc_rng_exponential": (f64, [vp, f64]),
            "orc_rng_uniform_int": (i32, [vp, i32, i32]),
            "orc_rng_ndraws": (u64, [vp]),
            "orc_philox4x32_10": (None, [vp, vp, vp]),
            "orc_neglog_u53": (f64, [u64]),
            "orc_eg_create": (vp, [C.POINTER(Opt), i32]),
            "orc_eg_destroy": (None, [vp]),
            "orc_eg_push": (i32, [vp, C.POINTER(Event)]),
            "orc_eg_pop": (i32, [vp, C.POINTER(Event)]),
            "orc_eg_reassign": (i32, [vp, i32, i32, i32, i32]),
            "orc_eg_generate_new": (i32, [vp, f64, i32, i32]),
            "orc_eg_generate_hoff_new": (i32, [vp, f64, i32, i32, i32]),
            "orc_eg_generate_end": (i32, [vp, f64, i32, i32, i32]),
            "orc_eg_generate_hoff_end": (i32, [vp, f64, i32, i32, i32]),
            "orc_eg_id": (i64, [vp]),
            "orc_eg_sizes": (i32, [vp, vp, vp, vp]),
            "orc_eg_ndraws": (u64, [vp]),
            "orc_sim_create": (vp, [C.POINTER(Opt)]),
            "orc_sim_destroy": (None, [vp]),
            "orc_sim_run": (i32, [vp, i64, vp, i64, vp]),
            "orc_sim_read": (None, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "orc_sim_period": (None, [vp, vp, vp, i32]),
            "orc_sim_ndraws": (u64, [vp]),
            "orc_grid_hash": (u64, [vp]),
            "orc_frep_hash": (u64, [vp]),
            "orc_stats_cumus": (None, [vp, vp]),
            "orc_run_replicas": (i64, [C.POINTER(Opt), i64, i32, i64, i32, vp, vp, vp]),
            "orc_max_threads": (i32, []),
            "cpf_create": (vp, [C.POINTER(Opt)]),
            "cpf_destroy": (None, [vp]),
            "cpf_run": (i32, [vp, i64, vp, i64, vp]),
            "cpf_read": (None, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
            "cpf_ndraws": (u64, [vp]),
            "cpf_run_replicas": (i64, [C.POINTER(Opt), i64, i32, i64, i32, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_opt(**kw) -> Opt:
    o = Opt()
    lib().orc_opt_defaults(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


# ---- neighbourhoods --------------------------------------------------------
def _cells(fn, *args):
    buf = np.zeros(2 * 64, dtype=np.int32)
    n = fn(*args, _p(buf))
    return [tuple(int(v) for v in buf[2 * i : 2 * i + 2]) for i in range(n)]


def periphery(d, r, c):
    return _cells(lib().orc_periphery, d, r, c)


def get_neighs(d, r, c, include_self):
    return _cells(lib().orc_get_neighs, d, r, c, int(include_self))


def get_neighborhood_acc(d, r, c, include_self):
    return _cells(lib().orc_get_neighborhood_acc, d, r, c, int(include_self))


def table_sizes():
    a, b, c2 = C.c_int(), C.c_int(), C.c_int()
    m = np.zeros((49, 5), dtype=np.int64)
    lib().orc_table_sizes(C.byref(a), C.byref(b), C.byref(c2), _p(m))
    return a.value, b.value, c2.value, m


# ---- grid functions --------------------------------------------------------
def mk_grid():
    return np.zeros((ROWS, COLS, CHANNELS), dtype=np.uint8)


def _grid(g):
    g = np.ascontiguousarray(g, dtype=np.uint8)
    assert g.size == GRID_SIZE
    return g


def _frep(f):
    f = np.ascontiguousarray(f, dtype=np.int64)
    assert f.size == FREP_SIZE
    return f


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.size == FREP_SIZE
    return a


def eligible_chs(grid, cell):
    out = np.zeros(CHANNELS, dtype=np.int64)
    n = lib().orc_eligible_chs(_p(_grid(grid)), cell[0], cell[1], _p(out))
    return out[:n].copy()


def inuse_chs(grid, cell):
    out = np.zeros(CHANNELS, dtype=np.int64)
    n = lib().orc_inuse_chs(_p(_grid(grid)), cell[0], cell[1], _p(out))
    return out[:n].copy()


def violates_reuse_constraint(grid) -> bool:
    return bool(lib().orc_violates_reuse_constraint(_p(_grid(grid))))


def afterstate(grid, cell, ch, val: bool):
    g = _grid(grid).copy().reshape(ROWS, COLS, CHANNELS)
    g[cell[0], cell[1], ch] = 1 if val else 0
    return g


def execute_action(is_end, cell, ch, grid):
    g = _grid(grid).copy().reshape(ROWS, COLS, CHANNELS)
    lib().orc_execute_action(int(is_end), cell[0], cell[1], int(ch), _p(g))
    return g


def afterstates(grid, cell, is_end, chs):
    chs = np.ascontiguousarray(chs, dtype=np.int64)
    out = np.zeros((len(chs), ROWS, COLS, CHANNELS), dtype=np.uint8)
    lib().orc_afterstates(_p(_grid(grid)), cell[0], cell[1], int(is_end), _p(chs), len(chs), _p(out))
    return out


def feature_rep(grid):
    out = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    lib().orc_feature_rep(_p(_grid(grid)), _p(out))
    return out


def feature_rep_slow(grid):
    out = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    lib().orc_feature_rep_slow(_p(_grid(grid)), _p(out))
    return out


def inc_afterstate_freps(cell, is_end, chs, grid, frep):
    chs = np.ascontiguousarray(chs, dtype=np.int64)
    out = np.zeros((len(chs), ROWS, COLS, NFEATS), dtype=np.int64)
    lib().orc_inc_afterstate_freps(
        cell[0], cell[1], int(is_end), _p(chs), len(chs), _p(_grid(grid)), _p(_frep(frep)), _p(out)
    )
    return out


def bool_sum(grid) -> int:
    return int(lib().orc_bool_sum(_p(_grid(grid))))


# ---- agent -----------------------------------------------------------------
def dense_dot(order, frep, w) -> np.float32:
    return np.float32(lib().orc_dense_dot(order, _p(_frep(frep)), _p(_f32(w))))


def dense_dot_g(order, frep, wg) -> np.float32:
    """x . wGradCorr (Agent.hs:66): the FAST order uses its own (thread-mapped) tree for this dot."""
    return np.float32(lib().orc_dense_dot_g(order, _p(_frep(frep)), _p(_f32(wg))))


def argpmax(q) -> int:
    q = np.ascontiguousarray(q, dtype=np.float32)
    return int(lib().orc_argpmax(_p(q), len(q)))


def select_action(order, cell, is_end, chs, w, grid, frep):
    chs = np.ascontiguousarray(chs, dtype=np.int64)
    qval = C.c_float()
    afrep = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    qvals = np.zeros(len(chs), dtype=np.float32)
    idx = lib().orc_select_action(
        order, cell[0], cell[1], int(is_end), _p(chs), len(chs), _p(_f32(w)), _p(_grid(grid)),
        _p(_frep(frep)), C.byref(qval), _p(afrep), _p(qvals),
    )
    return idx, np.float32(qval.value), afrep, qvals


def get_action(order, cell, is_end, w, grid, frep):
    nf = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    ncand = C.c_int()
    ch = lib().orc_get_action(
        order, cell[0], cell[1], int(is_end), _p(_f32(w)), _p(_grid(grid)), _p(_frep(frep)), _p(nf),
        C.byref(ncand),
    )
    return ch, nf, ncand.value


def get_action_lookahead(order, cell, hoff_cell, w, grid, frep):
    """hand-off look-ahead (extension): -> (ch, next_frep, ncand, q[ncand])"""
    nf = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
    ncand = C.c_int()
    q = np.zeros(CHANNELS, dtype=np.float32)
    ch = lib().orc_get_action_lookahead(order, cell[0], cell[1], hoff_cell[0], hoff_cell[1], _p(_f32(w)), _p(_grid(grid)),
                                        _p(_frep(frep)), _p(nf), C.byref(ncand), _p(q))
    return ch, nf, ncand.value, q[: ncand.value].copy()


def backward(order, aN, aA, aG, frep, next_frep, reward, w, wg, avg_reward):
    w = _f32(w).copy()
    wg = _f32(wg).copy()
    avg = C.c_float(avg_reward)
    isnan = C.c_int()
    loss = lib().orc_backward(
        order, aN, aA, aG, _p(_frep(frep)), _p(_frep(next_frep)), float(reward), _p(w), _p(wg),
        C.byref(avg), C.byref(isnan),
    )
    return np.float32(loss), bool(isnan.value), w, wg, np.float32(avg.value)


def grid_step_train(order, aN, aA, aG, is_end, cell, ch, frep, next_frep, grid, w, wg, avg_reward):
    grid = _grid(grid).copy().reshape(ROWS, COLS, CHANNELS)
    w = _f32(w).copy()
    wg = _f32(wg).copy()
    avg = C.c_float(avg_reward)
    isnan = C.c_int()
    reward = C.c_int64()
    loss = lib().orc_grid_step_train(
        order, aN, aA, aG, int(is_end), cell[0], cell[1], int(ch), _p(_frep(frep)), _p(_frep(next_frep)),
        _p(grid), _p(w), _p(wg), C.byref(avg), C.byref(isnan), C.byref(reward),
    )
    return grid, w, wg, np.float32(avg.value), np.float32(loss), bool(isnan.value), int(reward.value)


# ---- rng -------------------------------------------------------------------
class Rng:
    def __init__(self, mode=RNG_MT, seed=0, replica=0):
        self.h = lib().orc_rng_create(mode, seed, replica)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_rng_destroy(self.h)
            self.h = None

    def word64(self):
        return int(lib().orc_rng_word64(self.h))

    def double(self):
        return float(lib().orc_rng_double(self.h))

    def exponential(self, mean):
        return float(lib().orc_rng_exponential(self.h, mean))

    def uniform_int(self, lo, hi):
        return int(lib().orc_rng_uniform_int(self.h, lo, hi))

    @property
    def ndraws(self):
        return int(lib().orc_rng_ndraws(self.h))


def philox4x32_10(ctr, key):
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def neglog_u53(word: int) -> float:
    return float(lib().orc_neglog_u53(word))


# ---- eventgen --------------------------------------------------------------
class EventGen:
    def __init__(self, opt: Opt | None = None, with_initial_events=False):
        self.opt = opt or default_opt()
        self.h = lib().orc_eg_create(C.byref(self.opt), int(with_initial_events))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_eg_destroy(self.h)
            self.h = None

    def push(self, time, etype, cell, ch=-1, hoff=None):
        ev = Event(time, etype, cell[0], cell[1], ch, *(hoff if hoff else (-1, -1)))
        return lib().orc_eg_push(self.h, C.byref(ev))

    def pop(self):
        ev = Event()
        if lib().orc_eg_pop(self.h, C.byref(ev)) != 0:
            raise IndexError("pop from empty EventGen")
        return ev

    def reassign(self, cell, from_ch, to_ch):
        if lib().orc_eg_reassign(self.h, cell[0], cell[1], from_ch, to_ch) != 0:
            raise KeyError((cell, from_ch))

    def generate_new(self, time, cell):
        lib().orc_eg_generate_new(self.h, time, cell[0], cell[1])

    def generate_hoff_new(self, time, cell, ch):
        lib().orc_eg_generate_hoff_new(self.h, time, cell[0], cell[1], ch)

    def generate_end(self, time, cell, ch):
        lib().orc_eg_generate_end(self.h, time, cell[0], cell[1], ch)

    def generate_hoff_end(self, time, cell, ch):
        lib().orc_eg_generate_hoff_end(self.h, time, cell[0], cell[1], ch)

    @property
    def eg_id(self):
        return int(lib().orc_eg_id(self.h))

    def sizes(self):
        a, b, c2 = C.c_int(), C.c_int(), C.c_int()
        lib().orc_eg_sizes(self.h, C.byref(a), C.byref(b), C.byref(c2))
        return a.value, b.value, c2.value

    def is_empty(self):
        return self.sizes() == (0, 0, 0)

    @property
    def ndraws(self):
        return int(lib().orc_eg_ndraws(self.h))


# ---- simulator -------------------------------------------------------------
class Sim:
    def __init__(self, opt: Opt | None = None, **kw):
        self.opt = opt or default_opt(**kw)
        self.h = lib().orc_sim_create(C.byref(self.opt))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sim_destroy(self.h)
            self.h = None

    def run(self, max_steps, trace=False):
        tr = np.zeros(max_steps if trace else 0, dtype=TRACE_DTYPE)
        done = C.c_int64()
        status = lib().orc_sim_run(self.h, max_steps, _p(tr) if trace else None, len(tr), C.byref(done))
        return status, done.value, (tr[: done.value] if trace else None)

    def read(self):
        grid = np.zeros((ROWS, COLS, CHANNELS), dtype=np.uint8)
        frep = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
        w = np.zeros(FREP_SIZE, dtype=np.float32)
        wg = np.zeros(FREP_SIZE, dtype=np.float32)
        avg = C.c_float()
        it = C.c_int64()
        ev = Event()
        stats = np.zeros(8, dtype=np.int64)
        lib().orc_sim_read(self.h, _p(grid), _p(frep), _p(w), _p(wg), C.byref(avg), C.byref(it), C.byref(ev), _p(stats))
        return dict(grid=grid, frep=frep, w=w, wg=wg, avg_reward=np.float32(avg.value), iter=it.value,
                    event=ev.astuple(), stats=stats)

    def period(self, reset=True):
        s = C.c_float()
        n = C.c_int64()
        lib().orc_sim_period(self.h, C.byref(s), C.byref(n), int(reset))
        return np.float32(s.value), n.value

    @property
    def ndraws(self):
        return int(lib().orc_sim_ndraws(self.h))


class FastSim(Sim):
    """The optimised CPU competitor (oracle/dca_cpu_fast.c): FAST order, delta-form q-values, AVX2, same results
    as Sim(order=ORDER_FAST) bit for bit."""

    def __init__(self, opt: Opt | None = None, **kw):
        self.opt = opt or default_opt(**kw)
        self.opt.order = ORDER_FAST
        self.h = lib().cpf_create(C.byref(self.opt))

    def __del__(self):
        if getattr(self, "h", None):
            lib().cpf_destroy(self.h)
            self.h = None

    def run(self, max_steps, trace=False):
        tr = np.zeros(max_steps if trace else 0, dtype=TRACE_DTYPE)
        done = C.c_int64()
        status = lib().cpf_run(self.h, max_steps, _p(tr) if trace else None, len(tr), C.byref(done))
        return status, done.value, (tr[: done.value] if trace else None)

    def read(self):
        grid = np.zeros((ROWS, COLS, CHANNELS), dtype=np.uint8)
        frep = np.zeros((ROWS, COLS, NFEATS), dtype=np.int64)
        w = np.zeros(FREP_SIZE, dtype=np.float32)
        wg = np.zeros(FREP_SIZE, dtype=np.float32)
        avg, it, ev = C.c_float(), C.c_int64(), Event()
        stats = np.zeros(8, dtype=np.int64)
        lib().cpf_read(self.h, _p(grid), _p(frep), _p(w), _p(wg), C.byref(avg), C.byref(it), C.byref(ev), _p(stats))
        return dict(grid=grid, frep=frep, w=w, wg=wg, avg_reward=np.float32(avg.value), iter=it.value,
                    event=ev.astuple(), stats=stats)

    def period(self, reset=True):
        raise NotImplementedError

    @property
    def ndraws(self):
        return int(lib().cpf_ndraws(self.h))


def run_replicas_fast(opt: Opt, first_replica, n_replicas, n_events, n_threads=0):
    """OpenMP batch of the optimised CPU competitor; same return value as run_replicas."""
    stats = np.zeros(8, dtype=np.int64)
    avg = np.zeros(n_replicas, dtype=np.float32)
    bp = np.zeros(n_replicas, dtype=np.float64)
    total = lib().cpf_run_replicas(C.byref(opt), first_replica, n_replicas, n_events, n_threads, _p(stats), _p(avg), _p(bp))
    return int(total), stats, avg, bp


def grid_hash(grid) -> int:
    return int(lib().orc_grid_hash(_p(_grid(grid))))


def frep_hash(frep) -> int:
    return int(lib().orc_frep_hash(_p(_frep(frep))))


def stats_cumus(stats8):
    st = np.ascontiguousarray(stats8, dtype=np.int64)
    out = np.zeros(3, dtype=np.float64)
    lib().orc_stats_cumus(_p(st), _p(out))
    return out


def run_replicas(opt: Opt, first_replica, n_replicas, n_events, n_threads=0):
    stats = np.zeros(8, dtype=np.int64)
    avg = np.zeros(n_replicas, dtype=np.float32)
    bp = np.zeros(n_replicas, dtype=np.float64)
    total = lib().orc_run_replicas(C.byref(opt), first_replica, n_replicas, n_events, n_threads, _p(stats), _p(avg), _p(bp))
    return int(total), stats, avg, bp


def max_threads() -> int:
    return int(lib().orc_max_threads())
