"""runSim on the device (src/SimRunner.hs:206-243): periods of log_iter events, the reference's report
lines, and BASELINE config 5 (a call_rate x learning-rate grid of replicas with per-replica parameters)."""
import numpy as np
import pytest

import pyoracle as po
from haskelldca_b200 import Opt, Replicas, sim_runner, stats
from haskelldca_b200.replica_shards import per_point_blocking, slice_params, sweep_grid

pytestmark = pytest.mark.gpu


def test_run_sim_reports_like_the_reference():
    """--fixed_rng, 3000 events, log_iter 1000: every period line and the end report are what the oracle's
    counters give through the same format strings (Stats.hs:108-115, 130-151)."""
    n, li = 3000, 1000
    lines = []
    res = sim_runner.run_sim(Opt(n_events=n, log_iter=li, order="ref", rng="mt19937", rng_seed=0), out=lines.append)
    assert res["status"] == po.SUCCESS
    o = po.Sim(po.default_opt(order=po.ORDER_REF, rng_mode=po.RNG_MT, seed=0, n_events=n, log_iter=li))
    want = []
    for p in range(n // li):
        o.run(li)
        st = o.read()
        loss_sum, loss_n = o.period(reset=True)
        want.append(stats.stats_report_period(st["stats"], st["iter"], li, loss_sum, loss_n, st["avg_reward"]))
    st = o.read()
    assert lines[: n // li] == want
    assert lines[n // li] == "Success"
    assert lines[n // li + 1] == stats.stats_report_end(st["stats"], int(st["grid"].sum()), st["iter"])
    assert lines[0].startswith("Blocking probability events 0-1000: ") and "avg. reward" in lines[0]
    assert "events/second" in lines[-1]


def test_config5_sweep_grid_of_replicas():
    """call_rate x learning-rate grid as per-replica parameter arrays; two shards of the grid run as two
    handles (what two ranks would do) give the same per-point results as one handle; spot-check vs the oracle."""
    rates = [150.0, 225.0, 300.0]
    lrs = [(2.52e-6, 0.06, 5e-6), (1e-6, 0.03, 1e-6)]
    per, point, points = sweep_grid(rates, lrs, seeds_per_point=3)
    n_rep, n_ev = len(point), 1500
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=21), per_replica=per) as sims:
        sims.run(n_ev)
        whole = sims.read_summary()
    halves = []
    for first, count in ((0, 9), (9, 9)):
        with Replicas(Opt(n_replicas=count, n_events=n_ev, order="fast", rng="philox", rng_seed=21, first_replica=first),
                      per_replica=slice_params(per, first, count)) as sims:
            sims.run(n_ev)
            halves.append(sims.read_summary())
    for k in ("status", "iter", "stats"):
        assert np.array_equal(whole[k], np.concatenate([h[k] for h in halves])), k
    assert np.array_equal(whole["avg_reward"].view(np.uint32), np.concatenate([h["avg_reward"] for h in halves]).view(np.uint32))
    bp = per_point_blocking(whole["stats"], point, len(points))
    assert bp.shape == (6,) and bp[4:].mean() > bp[:2].mean()  # 300 calls/h blocks more than 150 calls/h
    for i in (0, 7, 17):
        cr, (an, aa, ag) = points[point[i]]
        o = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=21, replica=i, n_events=n_ev, call_rate=cr,
                                  alpha_net=an, alpha_avg=aa, alpha_grad=ag))
        o.run(n_ev)
        st = o.read()
        assert np.array_equal(st["stats"], whole["stats"][i]) and np.float32(st["avg_reward"]).view(np.uint32) == whole["avg_reward"][i].view(np.uint32)


def test_allreduce_stats_over_two_devices_in_one_process():
    """dca_allreduce_stats: one handle per device, ncclAllReduce of the int64[10] counters (the only collective
    of the path).  Needs two GPUs; skipped on a one-GPU box."""
    import ctypes as C

    from haskelldca_b200 import _ffi

    L = _ffi.load()
    if L.dca_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    n_rep, n_ev = 32, 1000
    handles, totals = [], []
    try:
        for dev in range(2):
            s = Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=4, device=dev,
                             first_replica=dev * n_rep, hoff_prob=0.1))
            s.run(n_ev)
            handles.append(s)
            totals.append(s.sum_stats()[0])
        arr = (C.c_void_p * 2)(handles[0].h, handles[1].h)
        out = np.zeros(10, dtype=np.int64)
        _ffi.check(L.dca_allreduce_stats(arr, 2, _ffi.ptr(out)), handles[0].h)
        assert np.array_equal(out, totals[0] + totals[1])
        assert out[8] == 2 * n_rep * n_ev and out[9] == 0
        # the same 64 replicas on one device give the same counters
        with Replicas(Opt(n_replicas=2 * n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=4, hoff_prob=0.1)) as one:
            one.run(n_ev)
            assert np.array_equal(one.sum_stats()[0], out)
    finally:
        for s in handles:
            s.close()
