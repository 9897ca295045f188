"""The optimised CPU competitor (oracle/dca_cpu_fast.c: delta-form q-values, bit-mask grid, AVX2, OpenMP) is
the SAME program as the literal oracle in ORDER_FAST: every trace record and the final state agree bit for bit.
That is what makes its events/s the fair CPU number in bench.py's cpu_baseline, next to
the literal dense port of the reference's formulation (src/Gridfuncs.hs:186-252, src/Simulator.hs:185-205)."""
import numpy as np
import pytest

import pyoracle as po


def _same_trace(ta, tb):
    assert len(ta) == len(tb)
    for k in ta.dtype.names:
        if k == "pad":
            continue
        x, y = ta[k], tb[k]
        if x.dtype.kind == "f":
            u = "u4" if x.dtype.itemsize == 4 else "u8"
            nan = np.isnan(x) & np.isnan(y)
            x, y = np.where(nan, 0, x.view(u)), np.where(nan, 0, y.view(u))
        bad = np.flatnonzero(x != y)
        assert bad.size == 0, f"{k} first differs at event {bad[0]}: {ta[k][bad[0]]} vs {tb[k][bad[0]]}"


def _same_state(a, b):
    ra, rb = a.read(), b.read()
    for k in ("grid", "frep", "stats"):
        assert np.array_equal(ra[k], rb[k]), k
    for k in ("w", "wg"):
        assert np.array_equal(ra[k].view(np.uint32), rb[k].view(np.uint32)), k
    assert np.float32(ra["avg_reward"]).view(np.uint32) == np.float32(rb["avg_reward"]).view(np.uint32)
    assert ra["iter"] == rb["iter"] and ra["event"] == rb["event"] and a.ndraws == b.ndraws


@pytest.mark.parametrize("kw", [
    dict(),                                                       # BASELINE config 1 / 3 parameters
    dict(hoff_prob=0.15),                                         # config 2: hand-offs
    dict(call_rate=1000.0),                                       # saturated grid, many events without a candidate
    dict(call_rate=40.0, call_dur=0.5, call_dur_hoff=0.2),        # nearly empty grid: ~70 candidates per arrival
    dict(call_rate=300.0, call_dur=10.0, call_dur_hoff=10.0, hoff_prob=0.9),
    dict(alpha_net=2e-5, alpha_avg=0.3, alpha_grad=5e-5),
])
@pytest.mark.parametrize("rng", [po.RNG_PHILOX, po.RNG_MT])
def test_optimised_cpu_path_equals_the_literal_oracle(kw, rng):
    n = 4000
    base = dict(rng_mode=rng, seed=11, replica=5, n_events=n, **kw)
    a = po.Sim(po.default_opt(order=po.ORDER_FAST, **base))
    b = po.FastSim(po.default_opt(**base))
    sa, da, ta = a.run(n, trace=True)
    sb, db, tb = b.run(n, trace=True)
    assert (sa, da) == (sb, db)
    _same_trace(ta, tb)
    _same_state(a, b)


def test_optimised_cpu_path_stop_rules_and_chunked_runs():
    n = 1500
    a = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=2, n_events=n, min_loss=1e9))
    b = po.FastSim(po.default_opt(rng_mode=po.RNG_PHILOX, seed=2, n_events=n, min_loss=1e9))
    assert a.run(n)[:2] == b.run(n)[:2] == (po.ZERO_LOSS, 51)
    a = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=2, n_events=n, hoff_prob=0.2))
    b = po.FastSim(po.default_opt(rng_mode=po.RNG_PHILOX, seed=2, n_events=n, hoff_prob=0.2))
    a.run(n)
    for chunk in (1, 99, 400, 1000, 50):
        b.run(chunk)
    _same_state(a, b)
    assert b.run(10)[:2] == (po.SUCCESS, 0)  # a finished simulator does not move


def test_optimised_batch_equals_literal_batch():
    o = po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=0, n_events=800)
    ta, sa, avga, bpa = po.run_replicas(o, 100, 6, 800, 2)
    tb, sb, avgb, bpb = po.run_replicas_fast(o, 100, 6, 800, 2)
    assert ta == tb == 4800 and np.array_equal(sa, sb)
    assert np.array_equal(avga.view(np.uint32), avgb.view(np.uint32)) and np.array_equal(bpa, bpb)
