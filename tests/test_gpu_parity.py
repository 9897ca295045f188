"""GPU parity: libdca_b200 (through the C ABI) vs the CPU oracle.  Run with -m gpu on a B200.

Bar: grid, frep, chosen channels, event stream bit-exact; weights, average reward and loss
bit-exact too (the oracle fixes the fp32 order per mode) -- a fortiori within the 1e-4
relative tolerance BASELINE.json's north_star states.
"""
import numpy as np
import pytest

import pyoracle as po

pytestmark = pytest.mark.gpu

TRACE_EXACT = ("etype", "cell", "ch", "ncand", "end_ch", "reward", "time", "grid_hash", "frep_hash")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(a, b):
    """Bit equality of fp32 arrays; NaN payloads are not part of the contract (x86 makes 0xffc00000, the GPU
    0x7fffffff), so NaN == NaN here."""
    a, b = np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


def assert_trace_equal(got, want):
    assert len(got) == len(want)
    for k in TRACE_EXACT:
        bad = np.flatnonzero(got[k] != want[k])
        assert bad.size == 0, f"trace field {k} first differs at event {bad[0]}: {got[k][bad[0]]} vs {want[k][bad[0]]}"
    for k in ("loss", "avg_reward"):
        bad = np.flatnonzero(~same_bits(got[k], want[k]))
        assert bad.size == 0, f"trace field {k} first differs at event {bad[0]}: {got[k][bad[0]]} vs {want[k][bad[0]]}"


def assert_state_equal(got, want, rel=None):
    for k in ("grid", "frep", "stats"):
        assert np.array_equal(got[k], want[k]), k
    assert got["iter"] == want["iter"]
    assert got["event"] == want["event"]
    for k in ("w", "wg"):
        assert same_bits(got[k], want[k]).all(), k
        np.testing.assert_allclose(got[k], want[k], rtol=1e-4)  # the stated fp32 tolerance, implied by bit equality
    assert same_bits(got["avg_reward"], want["avg_reward"]).all()


def oracle_run(order, n_events, seed=0, hoff=0.0, rng=po.RNG_MT, replica=0, **kw):
    o = po.Sim(po.default_opt(order=order, rng_mode=rng, seed=seed, replica=replica, n_events=n_events, hoff_prob=hoff, **kw))
    status, done, tr = o.run(n_events, trace=True)
    return o, status, tr


@pytest.mark.parametrize("order,oorder", [("ref", po.ORDER_REF), ("fast", po.ORDER_FAST)])
def test_config1_fixed_rng_replay_10k(order, oorder):
    """BASELINE config 1: 7x7x70, 200/h, 3 min, hoff 0, 10 000 events, --fixed_rng (MT19937-64 seed 0)."""
    from haskelldca_b200 import Opt, Replicas

    n = 10000
    o, status, tr = oracle_run(oorder, n)
    with Replicas(Opt(n_events=n, order=order, rng="mt19937", rng_seed=0, trace=n)) as sims:
        sims.run(n)
        assert sims.read_status(0)[0] == status == po.SUCCESS
        assert sims.read_status(0)[2] == o.ndraws
        assert_trace_equal(sims.read_trace(0), tr)
        got, want = sims.read_state(0), o.read()
        assert_state_equal(got, want)
        bp_got = got["stats"][2] / (got["stats"][0] + 1.0)
        assert bp_got == pytest.approx(po.stats_cumus(want["stats"])[0], rel=1e-4)


@pytest.mark.parametrize("order,oorder,n", [("ref", po.ORDER_REF, 100000), ("fast", po.ORDER_FAST, 100000)])
def test_config2_handoffs(order, oorder, n):
    """BASELINE config 2: hoff_prob 0.15, call_dur_hoff 1, 100 000 events, --fixed_rng, both arithmetic orders."""
    from haskelldca_b200 import Opt, Replicas

    o, status, tr = oracle_run(oorder, n, hoff=0.15)
    with Replicas(Opt(n_events=n, order=order, rng="mt19937", rng_seed=0, hoff_prob=0.15, trace=n)) as sims:
        sims.run(n)
        assert sims.read_status(0)[0] == status
        assert_trace_equal(sims.read_trace(0), tr)
        assert_state_equal(sims.read_state(0), o.read())
    assert (tr["etype"] == po.HOFF).sum() > 500


def test_run_in_periods_equals_one_run():
    from haskelldca_b200 import Opt, Replicas

    n = 3000
    o, _, tr = oracle_run(po.ORDER_FAST, n, hoff=0.1)
    with Replicas(Opt(n_events=n, order="fast", rng="mt19937", rng_seed=0, hoff_prob=0.1, trace=n)) as sims:
        for _ in range(6):
            sims.run(500)  # runPeriod-sized chunks (SimRunner.hs:167-168)
        assert sims.read_status(0)[0] == po.SUCCESS
        assert_trace_equal(sims.read_trace(0), tr)
        assert_state_equal(sims.read_state(0), o.read())
        sims.run(500)  # a finished replica does not move
        assert sims.read_status(0)[1] == n


@pytest.mark.parametrize("first,seed", [(0, 42), (1000, 42), ((1 << 33) + 5, (7 << 32) | 3)])
def test_philox_replicas_match_oracle(first, seed):
    """Config 3's RNG mode on a small batch: every replica equals its own oracle run; the Philox stream id
    (first_replica + r) and the key (seed) are 64-bit quantities."""
    from haskelldca_b200 import Opt, Replicas

    n_rep, n_ev = 24, 1500
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=seed, hoff_prob=0.15,
                      first_replica=first, trace=n_ev)) as sims:
        sims.run(n_ev)
        summ = sims.read_summary()
        assert (summ["status"] == po.SUCCESS).all() and (summ["iter"] == n_ev).all()
        for r in range(n_rep):
            o, _, tr = oracle_run(po.ORDER_FAST, n_ev, seed=seed, hoff=0.15, rng=po.RNG_PHILOX, replica=first + r)
            assert_trace_equal(sims.read_trace(r), tr)
            assert_state_equal(sims.read_state(r), o.read())
        total, dev = sims.sum_stats()
        assert total[8] == n_rep * n_ev and total[9] == 0 and dev
        assert np.array_equal(total[:8], summ["stats"].sum(axis=0))


def test_philox_ref_order_matches_oracle():
    from haskelldca_b200 import Opt, Replicas

    n_rep, n_ev = 4, 600
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="ref", rng="philox", rng_seed=5, hoff_prob=0.2, trace=n_ev)) as sims:
        sims.run(n_ev)
        for r in range(n_rep):
            o, _, tr = oracle_run(po.ORDER_REF, n_ev, seed=5, hoff=0.2, rng=po.RNG_PHILOX, replica=r)
            assert_trace_equal(sims.read_trace(r), tr)
            assert_state_equal(sims.read_state(r), o.read())


def test_per_replica_parameters():
    """Config 5's shape: call_rate x learning-rate grid as per-replica arrays."""
    from haskelldca_b200 import Opt, Replicas

    rates = [150.0, 200.0, 250.0, 300.0]
    lrs = [(2.52e-6, 0.06, 5e-6), (1e-6, 0.03, 1e-6)]
    grid = [(r, lr) for r in rates for lr in lrs]
    n_ev = 1200
    per = dict(call_rate=[g[0] for g in grid], learning_rate=[g[1][0] for g in grid],
               learning_rate_avg=[g[1][1] for g in grid], learning_rate_grad=[g[1][2] for g in grid])
    with Replicas(Opt(n_replicas=len(grid), n_events=n_ev, order="fast", rng="philox", rng_seed=9), per_replica=per) as sims:
        sims.run(n_ev)
        for i, (rate, lr) in enumerate(grid):
            o = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=9, replica=i, n_events=n_ev,
                                      call_rate=rate, alpha_net=lr[0], alpha_avg=lr[1], alpha_grad=lr[2]))
            o.run(n_ev)
            assert_state_equal(sims.read_state(i), o.read())
        # set_params + reset: all replicas at 300/h
        nb = sims.set_params(call_rate=[300.0] * len(grid))
        assert nb == 8 * len(grid)
        sims.reset()
        sims.run(n_ev)
        o = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=9, replica=1, n_events=n_ev,
                                  call_rate=300.0, alpha_net=grid[1][1][0], alpha_avg=grid[1][1][1], alpha_grad=grid[1][1][2]))
        o.run(n_ev)
        assert_state_equal(sims.read_state(1), o.read())


def test_reset_reproduces_the_run():
    from haskelldca_b200 import Opt, Replicas

    with Replicas(Opt(n_replicas=16, n_events=800, order="fast", rng="philox", rng_seed=3)) as sims:
        sims.run(800)
        a = sims.read_summary()
        w0 = sims.read_state(5)["w"].copy()
        sims.reset()
        assert (sims.read_summary()["iter"] == 0).all()
        sims.run(800)
        b = sims.read_summary()
        assert np.array_equal(a["stats"], b["stats"]) and np.array_equal(bits(a["avg_reward"]), bits(b["avg_reward"]))
        assert np.array_equal(bits(w0), bits(sims.read_state(5)["w"]))


def test_checkpoint_resume_continues_bit_for_bit():
    from haskelldca_b200 import Opt, Replicas
    from haskelldca_b200._ffi import DcaError

    opt = dict(n_replicas=6, n_events=1200, order="fast", rng="philox", rng_seed=17, hoff_prob=0.1)
    with Replicas(Opt(**opt)) as a:
        a.run(500)
        blob = a.checkpoint()
        a.run(700)
        want = [a.read_state(r) for r in range(6)]
        summ = a.read_summary()
    with Replicas(Opt(**opt)) as b:
        b.restore(blob)
        assert (b.read_summary()["iter"] == 500).all()
        b.run(700)
        for r in range(6):
            assert_state_equal(b.read_state(r), want[r])
        assert np.array_equal(b.read_summary()["stats"], summ["stats"])
    with Replicas(Opt(**{**opt, "n_replicas": 5})) as c, pytest.raises(DcaError):
        c.restore(blob)


def test_stepwise_operators_follow_the_oracle_event_stream():
    """accFn1/accFn2 through dca_get_action / dca_grid_step_train with the host keeping the event generator
    (the reference's runStep shape, SimRunner.hs:60-97)."""
    from haskelldca_b200 import Opt, Replicas

    for order, oorder in (("ref", po.ORDER_REF), ("fast", po.ORDER_FAST)):
        n = 250
        o, _, tr = oracle_run(oorder, n, hoff=0.15)
        with Replicas(Opt(n_events=n, order=order, rng="philox")) as sims:  # the device queue is not used here
            for i in range(n):
                cell = (int(tr["cell"][i]) // 7, int(tr["cell"][i]) % 7)
                is_end = tr["etype"][i] == po.END
                ch = sims.get_action(0, cell, is_end)
                assert ch == tr["ch"][i], (order, i)
                loss, isnan, reward = sims.grid_step_train(0, is_end, cell, ch)
                assert bits(loss) == bits(tr["loss"][i]) and not isnan and reward == tr["reward"][i], (order, i)
            got, want = sims.read_state(0), o.read()
            for k in ("grid", "frep"):
                assert np.array_equal(got[k], want[k])
            for k in ("w", "wg"):
                assert np.array_equal(bits(got[k]), bits(want[k]))
            assert bits(got["avg_reward"]) == bits(want["avg_reward"])


def test_stop_rules_and_status_words():
    from haskelldca_b200 import Opt, Replicas, _ffi

    with Replicas(Opt(n_replicas=3, n_events=1000, min_loss=1e9, order="fast")) as sims:  # ZeroLoss, SimRunner.hs:161
        sims.run(1000)
        s = sims.read_summary()
        assert (s["status"] == _ffi.ZERO_LOSS).all() and (s["iter"] == 51).all()
    with Replicas(Opt(n_replicas=2, n_events=100, order="fast", verify_reuse_constraint=True, verify_nelig_bounds=True)) as sims:
        sims.run(1000)
        s = sims.read_summary()
        assert (s["status"] == _ffi.SUCCESS).all() and (s["iter"] == 100).all()
    with Replicas(Opt(n_replicas=2, n_events=500, order="fast")) as sims:  # NaNLoss, Agent.hs:80 / SimRunner.hs:152
        sims.run(100)
        w = sims.read_state(1)["w"]
        w[:] = np.nan
        sims.write_state(1, w=w)
        sims.run(400)
        s = sims.read_summary()
        assert s["status"].tolist() == [_ffi.SUCCESS, _ffi.NAN_LOSS] and s["iter"].tolist() == [500, 101]
        assert sims.sum_stats()[0][9] == 1
    with Replicas(Opt(n_events=5000, order="fast", rng="mt19937")) as sims:  # tape shorter than the run
        sims.set_rng_mt19937(0, 0, 600)
        sims.run(5000)
        st, it, nd = sims.read_status(0)
        assert st == _ffi.TAPE_EXHAUSTED and 100 < it < 600 and nd <= 600
    # a corrupted grid trips the reuse-constraint check (SimRunner.hs:155-156)
    with Replicas(Opt(n_events=50, order="fast", verify_reuse_constraint=True)) as sims:
        g = np.zeros((7, 7, 70), np.uint8)
        g[3, 3, 5] = g[3, 4, 5] = 1
        sims.write_state(0, grid=g)
        sims.run(10)
        assert sims.read_status(0) [0] == _ffi.REUSE_VIOLATED and sims.read_status(0)[1] == 1


def test_period_accumulators():
    from haskelldca_b200 import Opt, Replicas

    n = 1000
    o, _, tr = oracle_run(po.ORDER_FAST, n)
    with Replicas(Opt(n_events=3000, order="fast", rng="mt19937", log_iter=n)) as sims:
        sims.run(n)
        lsum, ln = sims.read_period(0, reset=True)
        want_sum, want_n = o.period(reset=False)
        assert ln == want_n == n and bits(lsum) == bits(want_sum)
        st = sims.read_stats(0)
        assert st[6] == 0 and st[7] == 0 and st[0] == o.read()["stats"][0]
        assert sims.read_period(0)[1] == 0


def test_invariants_at_scale():
    """Size-independent properties on a larger batch (Stats.hs:139-147 call conservation, reuse constraint,
    incremental frep == featureRep, event counts)."""
    from haskelldca_b200 import Opt, Replicas, gridfuncs

    n_rep, n_ev = 4096, 2000  # BASELINE configs[2]'s replica count
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=11, hoff_prob=0.15)) as sims:
        ms = sims.run(n_ev)
        s = sims.read_summary()
        assert (s["status"] == po.SUCCESS).all() and (s["iter"] == n_ev).all() and ms > 0
        st = s["stats"]
        in_progress = st[:, 0] + st[:, 3] - st[:, 2] - st[:, 5] - st[:, 4]
        assert (st[:, 5] == 0).all() and (st[:, 0] + st[:, 3] + st[:, 4] == n_ev).all()
        assert len(np.unique(s["avg_reward"])) > n_rep // 2
        sims_w = {}
        for r in (0, 1, 2047, 4095):
            x = sims.read_state(r)
            sims_w[r] = x["w"].copy()
            assert x["grid"].sum() == in_progress[r]
            assert not gridfuncs.violates_reuse_constraint(x["grid"])
            assert np.array_equal(x["frep"], po.feature_rep(x["grid"]))
            assert np.array_equal(gridfuncs.feature_rep(x["grid"]), x["frep"])
        bp = st[:, 2].sum() / (st[:, 0].sum() + 1.0)
        assert 0.02 < bp < 0.4
        # the warp-specialised kernel is deterministic: a second run of the whole batch is bit-identical
        sims.reset()
        sims.run(n_ev)
        s2 = sims.read_summary()
        for k in ("status", "iter", "stats"):
            assert np.array_equal(s[k], s2[k]), k
        assert np.array_equal(s["avg_reward"].view(np.uint32), s2["avg_reward"].view(np.uint32))
        for r in (0, 4095):
            y = sims.read_state(r)
            assert np.array_equal(y["w"].view(np.uint32), sims_w[r].view(np.uint32))


def test_fast_and_ref_orders_agree_under_the_same_decisions():
    """The two arithmetic orders are different roundings of the same update.  Free-running they pick different
    channels after the first near-tie, so they are compared under the SAME event / action
    stream: the REF run's decisions are replayed through accFn2 in FAST order.  After 10 000 events weights and
    average reward must agree within north_star's fp32 tolerance of 1e-4 relative (measured on B200: 8.4e-6 on
    the weights, 1.8e-5 on the average reward)."""
    from haskelldca_b200 import Opt, Replicas

    n = 10000
    with Replicas(Opt(n_events=n, order="ref", rng="mt19937", rng_seed=0, trace=n)) as ref:
        ref.run(n)
        tr = ref.read_trace(0)
        want = ref.read_state(0)
    with Replicas(Opt(n_events=n, order="fast", rng="mt19937", rng_seed=0)) as fast:
        for e in tr:
            cell = (int(e["cell"]) // 7, int(e["cell"]) % 7)
            fast.grid_step_train(0, e["etype"] == po.END, cell, int(e["ch"]))
        got = fast.read_state(0)
    assert np.array_equal(got["grid"], want["grid"]) and np.array_equal(got["frep"], want["frep"])
    scale = np.abs(want["w"]).max()
    rel_w = np.abs(got["w"] - want["w"]).max() / scale
    rel_avg = abs(float(got["avg_reward"]) - float(want["avg_reward"])) / abs(float(want["avg_reward"]))
    print(f"FAST vs REF under forced actions after {n} events: max |dw|/max|w| = {rel_w:.2e}, avg reward rel = {rel_avg:.2e}")
    assert rel_w < 1e-4 and rel_avg < 1e-4  # the tolerance BASELINE.json's north_star states


@pytest.mark.parametrize("call_rate,call_dur,call_dur_hoff,hoff", [
    (1000.0, 3.0, 1.0, 0.0),    # saturated: most arrivals are rejected, many events without a candidate
    (40.0, 0.5, 0.2, 0.0),      # nearly empty grid: ~70 candidates on every arrival
    (250.0, 6.0, 0.5, 0.5),     # half of the calls are handed off
    (300.0, 10.0, 10.0, 0.9),   # long calls, almost always handed off: END+HOFF pairs dominate
])
def test_extreme_traffic_parameters_match_the_oracle(call_rate, call_dur, call_dur_hoff, hoff):
    from haskelldca_b200 import Opt, Replicas

    n_rep, n_ev = 3, 2500
    kw = dict(call_rate=call_rate, call_dur=call_dur, call_dur_hoff=call_dur_hoff, hoff_prob=hoff)
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=77, trace=n_ev, **kw)) as sims:
        sims.run(n_ev)
        for r in range(n_rep):
            o = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=77, replica=r, n_events=n_ev, **kw))
            status, done, tr = o.run(n_ev, trace=True)
            assert done == n_ev and sims.read_status(r)[0] == status == po.SUCCESS
            assert_trace_equal(sims.read_trace(r), tr)
            assert_state_equal(sims.read_state(r), o.read())


def test_randomised_parameters_and_chunked_runs_match_the_oracle():
    """Differential test over random traffic / learning parameters; the device run is cut into chunks of
    random length (every chunk is a separate launch that stages the replica in and out)."""
    from haskelldca_b200 import Opt, Replicas

    rng = np.random.default_rng(2026)
    for trial in range(24):
        order = ("fast", "ref")[trial % 3 == 2]
        oorder = po.ORDER_FAST if order == "fast" else po.ORDER_REF
        kw = dict(call_rate=float(rng.uniform(60, 600)), call_dur=float(rng.uniform(0.3, 8.0)),
                  call_dur_hoff=float(rng.uniform(0.2, 4.0)), hoff_prob=float(rng.choice([0.0, 0.05, 0.3, 0.7])))
        lr = dict(alpha_net=float(rng.choice([1e-7, 2.52e-6, 2e-5])), alpha_avg=float(rng.choice([0.01, 0.06, 0.3])),
                  alpha_grad=float(rng.choice([1e-6, 5e-6, 5e-5])))
        seed, n_ev = int(rng.integers(1, 1 << 30)), 900
        with Replicas(Opt(n_replicas=2, n_events=n_ev, order=order, rng="philox", rng_seed=seed, trace=n_ev, learning_rate=lr["alpha_net"],
                          learning_rate_avg=lr["alpha_avg"], learning_rate_grad=lr["alpha_grad"], **kw)) as sims:
            done = 0
            while done < n_ev:
                step = int(rng.integers(1, 400))
                sims.run(step)
                done += step
            for r in range(2):
                o = po.Sim(po.default_opt(order=oorder, rng_mode=po.RNG_PHILOX, seed=seed, replica=r, n_events=n_ev, **kw, **lr))
                status, n_done, tr = o.run(n_ev, trace=True)
                # a replica may legitimately stop early (zero / NaN loss); both sides must stop at the same event
                assert sims.read_status(r)[0] == status, (trial, kw, lr, status, n_done)
                assert len(tr) == n_done
                assert_trace_equal(sims.read_trace(r), tr)
                assert_state_equal(sims.read_state(r), o.read())


def test_guard_rails_of_the_c_abi():
    """Library faults are errors or dedicated status words, never a silent spin or an out-of-bounds write."""
    from haskelldca_b200 import Opt, Replicas, _ffi
    from haskelldca_b200._ffi import DcaError

    # a DCA_RNG_TAPE replica without a tape is DCA_UNINITIALIZED and dca_run refuses to start
    L = _ffi.load()
    import ctypes as C
    cfg = _ffi.Config()
    _ffi.check(L.dca_config_defaults(C.byref(cfg)))
    cfg.n_replicas, cfg.rng_mode, cfg.n_events = 2, _ffi.RNG_TAPE, 100
    h = C.c_void_p()
    _ffi.check(L.dca_create(C.byref(cfg), C.byref(h)))
    try:
        st = np.zeros(2, np.int32)
        _ffi.check(L.dca_read_summary(h, _ffi.ptr(st), None, None, None), h)
        assert (st == _ffi.UNINITIALIZED).all()
        _ffi.check(L.dca_set_rng_mt19937(h, 0, 0, 1000), h)  # replica 1 still has none
        assert L.dca_run(h, 10) == -4  # DCA_ERR_STATE
        assert b"replica 1 has no random tape" in L.dca_last_error(h)
        _ffi.check(L.dca_set_rng_mt19937(h, 1, 1, 1000), h)
        _ffi.check(L.dca_run(h, 10), h)
        _ffi.check(L.dca_read_summary(h, _ffi.ptr(st), None, None, None), h)
        assert (st == _ffi.RUNNING).all()
    finally:
        L.dca_destroy(h)

    # gridStepTrain only executes a channel getAction could have returned; a bad one changes nothing
    for order in ("fast", "ref"):
        with Replicas(Opt(n_events=100, order=order)) as sims:
            ch = sims.get_action(0, (3, 3), False)
            sims.grid_step_train(0, False, (3, 3), ch)
            before = sims.read_state(0)
            with pytest.raises(DcaError):
                sims.grid_step_train(0, False, (3, 4), ch)      # in use one cell away: not eligible
            with pytest.raises(DcaError):
                sims.grid_step_train(0, True, (3, 3), (ch + 1) % 70)  # not in use at the cell
            after = sims.read_state(0)
            assert np.array_equal(before["grid"], after["grid"]) and np.array_equal(bits(before["w"]), bits(after["w"]))
            sims.grid_step_train(0, True, (3, 3), ch)            # the legal END still works
            assert sims.read_state(0)["grid"].sum() == 0

    # a checkpoint only continues under the random stream it was saved with
    opt = dict(n_replicas=3, n_events=500, order="fast", rng="philox", rng_seed=17)
    with Replicas(Opt(**opt)) as a:
        a.run(100)
        blob = a.checkpoint()
    for other in (dict(rng_seed=18), dict(first_replica=3)):
        with Replicas(Opt(**{**opt, **other})) as b, pytest.raises(DcaError):
            b.restore(blob)
    with Replicas(Opt(n_events=500, order="fast", rng="mt19937", rng_seed=0)) as a:
        a.run(100)
        blob = a.checkpoint()
    with Replicas(Opt(n_events=500, order="fast", rng="mt19937", rng_seed=1)) as b, pytest.raises(DcaError):
        b.restore(blob)  # other tape
    with Replicas(Opt(n_events=500, order="fast", rng="mt19937", rng_seed=0)) as b:
        b.restore(blob)
        b.run(400)
        assert b.read_status(0)[:2] == (po.SUCCESS, 500)

    # more END events than the queue has slots (only reachable through a forced grid): a status word, not a crash
    for order in ("fast", "ref"):
        with Replicas(Opt(n_replicas=2, n_events=20000, order=order, call_rate=3000.0, call_dur=1e7)) as sims:
            for _ in range(4):  # every erase leaves the END events of the erased calls in the queue
                sims.run(2500)
                if sims.read_summary()["status"][1] != _ffi.RUNNING:
                    break
                assert sims.read_state(1)["grid"].sum() > 200
                sims.write_state(1, grid=np.zeros((7, 7, 70), np.uint8))
            s = sims.read_summary()
            assert s["status"][0] == _ffi.RUNNING and s["status"][1] == _ffi.QUEUE_OVERFLOW, s["status"]
            assert sims.sum_stats()[0][9] == 1


@pytest.mark.parametrize("order", ["ref", "fast"])
@pytest.mark.parametrize("rng,kw", [("philox", dict(hoff_prob=0.3)), ("mt19937", dict(hoff_prob=0.15, call_dur_hoff=1.0)),
                                    ("philox", dict(hoff_prob=0.6, call_rate=400.0, call_dur=1.0))])
def test_handoff_lookahead_matches_the_oracle(rng, kw, order):
    """The extension (the reference's TODO, README.md:71-72; include/dca_b200.h dca_config.hoff_lookahead) in the
    REF-order run kernel and in the fused FAST kernel (its own instantiation, k_run_fast<true, true>) against the
    oracle's first-principles definition: every event bit for bit."""
    from haskelldca_b200 import Opt, Replicas

    n_rep, n_ev = (3, 4000) if rng == "philox" else (1, 6000)
    orng = po.RNG_PHILOX if rng == "philox" else po.RNG_MT
    oorder = po.ORDER_REF if order == "ref" else po.ORDER_FAST
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order=order, rng=rng, rng_seed=0 if rng == "mt19937" else 12,
                      hoff_lookahead=True, trace=n_ev, **kw)) as sims:
        sims.run(n_ev)
        for r in range(n_rep):
            o = po.Sim(po.default_opt(order=oorder, rng_mode=orng, seed=0 if rng == "mt19937" else 12, replica=r, n_events=n_ev,
                                      hoff_lookahead=1, **kw))
            status, done, tr = o.run(n_ev, trace=True)
            assert sims.read_status(r)[0] == status and done == n_ev
            assert_trace_equal(sims.read_trace(r), tr)
            assert_state_equal(sims.read_state(r), o.read())
            # the look-ahead did steer some decisions: the plain run of the same stream differs
            if r == 0:
                p = po.Sim(po.default_opt(order=oorder, rng_mode=orng, seed=0 if rng == "mt19937" else 12, replica=r,
                                          n_events=n_ev, **kw))
                _, _, trp = p.run(n_ev, trace=True)
                assert not np.array_equal(trp["ch"], tr["ch"])


def test_handoff_lookahead_fast_without_trace_and_in_chunks():
    """The FAST look-ahead instantiation without a trace buffer, across unit boundaries (two launches, the second one
    longer than a unit of dca_run_unit_events()) and on many replicas: final states equal the oracle's."""
    from haskelldca_b200 import Opt, Replicas

    n_rep, n_ev, kw = 40, 9000, dict(hoff_prob=0.4, call_rate=300.0)
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=5, hoff_lookahead=True, **kw)) as sims:
        sims.run(3000)
        sims.run(6000)
        assert (sims.read_summary()["status"] == po.SUCCESS).all()
        for r in (0, 17, 39):
            o = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=5, replica=r, n_events=n_ev,
                                      hoff_lookahead=1, **kw))
            status, done, _ = o.run(n_ev)
            assert status == po.SUCCESS and done == n_ev
            assert_state_equal(sims.read_state(r), o.read())


def test_units_of_a_launch_chain_correctly():
    """dca_run (FAST order) cuts a launch into units of 4096 events per replica, one CTA each, chained through a
    release / acquire counter: a replica that stops in its first unit is skipped by its later ones, a launch whose
    length is not a multiple of the unit equals the oracle, and very many short units keep their order."""
    from haskelldca_b200 import Opt, Replicas, _ffi

    with Replicas(Opt(n_replicas=3, n_events=100000, min_loss=1e9, order="fast")) as sims:  # ZeroLoss at event 51
        sims.run(50000)  # 13 units per replica
        s = sims.read_summary()
        assert (s["status"] == _ffi.ZERO_LOSS).all() and (s["iter"] == 51).all()
    n = 4096 * 3 + 77
    o, status, tr = oracle_run(po.ORDER_FAST, n, seed=21, hoff=0.15, rng=po.RNG_PHILOX, replica=9)
    with Replicas(Opt(n_replicas=1, n_events=n, order="fast", rng="philox", rng_seed=21, hoff_prob=0.15, first_replica=9, trace=n)) as sims:
        sims.run(n)
        assert sims.read_status(0)[0] == status == po.SUCCESS
        assert_trace_equal(sims.read_trace(0), tr)
        assert_state_equal(sims.read_state(0), o.read())
    # one replica, 40 units in a single launch: every unit waits for its predecessor
    n = 4096 * 40
    o = po.FastSim(po.default_opt(rng_mode=po.RNG_PHILOX, seed=5, replica=0, n_events=n))
    o.run(n)
    with Replicas(Opt(n_replicas=1, n_events=n, order="fast", rng="philox", rng_seed=5)) as sims:
        sims.run(n)
        assert_state_equal(sims.read_state(0), o.read())


def test_async_run_and_run_many_equal_the_blocking_run():
    """dca_run_async / dca_wait / dca_run_many (one host thread, several handles in flight -- a handle per GPU of the
    box when there are several, two handles on one device otherwise): the same states as blocking dca_run calls, and
    any other call on a handle collects its launch in flight first."""
    import haskelldca_b200 as hd
    from haskelldca_b200 import Opt, Replicas, _ffi

    ndev = _ffi.load().dca_device_count()
    kw = [dict(n_replicas=3, n_events=5000, order="fast", rng="philox", rng_seed=3, first_replica=0, device=0),
          dict(n_replicas=2, n_events=5000, order="fast", rng="philox", rng_seed=3, first_replica=3, hoff_prob=0.15,
               device=1 if ndev > 1 else 0),
          dict(n_replicas=1, n_events=700, order="ref", rng="philox", rng_seed=4, device=0)]
    want = []
    for k in kw:
        with Replicas(Opt(**k)) as s:
            s.run(2000)
            s.run(k["n_events"])
            want.append([s.read_state(r) for r in range(k["n_replicas"])])
    sets = [Replicas(Opt(**k)) for k in kw]
    try:
        ms = hd.run_many(sets, 2000)
        assert len(ms) == 3 and all(m > 0 for m in ms)
        for s, k in zip(sets, kw):
            s.run_async(k["n_events"])
        # (no wait: reading a state collects the launch)
        for s, k, w in zip(sets, kw, want):
            for r in range(k["n_replicas"]):
                assert_state_equal(s.read_state(r), w[r])
            assert s.wait() > 0  # nothing in flight any more: returns the time of the collected launch
            assert (s.read_summary()["status"] == _ffi.SUCCESS).all()
    finally:
        for s in sets:
            s.close()


def test_ref_order_run_kernel_without_trace():
    """k_run<ORDER_REF, false> -- the instantiation a `RefOrder` user of dca_run (and bench.py's reference_order leg) gets:
    all folds of an event in lock step on one warp, wGradCorr in global memory, the event-sequential tail next to the
    dense update (run_events_overlapped).  Every other REF test runs the TRACE instantiation, which keeps the plain
    sequence.  Chunked launches (the state crosses global memory in between), hand-offs, the 70-candidate transient of
    the empty grid (folds on three warps), more replicas than fit the GPU at once."""
    from haskelldca_b200 import Opt, Replicas, _ffi

    n_rep, n_ev = 600, 420  # 444 CTAs of this kernel are resident on a 148-SM B200
    kw = dict(hoff_prob=0.2, call_dur_hoff=1.0)
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="ref", rng="philox", rng_seed=11, **kw)) as sims:
        for part in (1, 70, 149, 200):  # 420 events in four launches
            sims.run(part)
        s = sims.read_summary()
        assert (s["status"] == _ffi.SUCCESS).all() and (s["iter"] == n_ev).all()
        for r in (0, 1, 295, 443, 444, n_rep - 1):
            o = po.Sim(po.default_opt(order=po.ORDER_REF, rng_mode=po.RNG_PHILOX, seed=11, replica=r, n_events=n_ev,
                                      hoff_prob=0.2, call_dur_hoff=1.0))
            o.run(n_ev)
            assert_state_equal(sims.read_state(r), o.read())
    # look-ahead, verify flags and the stop rules on this instantiation
    with Replicas(Opt(n_replicas=3, n_events=900, order="ref", rng="philox", rng_seed=5, hoff_prob=0.3, hoff_lookahead=True,
                      verify_reuse_constraint=True, verify_nelig_bounds=True)) as sims:
        sims.run(900)
        for r in range(3):
            o = po.Sim(po.default_opt(order=po.ORDER_REF, rng_mode=po.RNG_PHILOX, seed=5, replica=r, n_events=900, hoff_prob=0.3,
                                      hoff_lookahead=1))
            o.run(900)
            assert_state_equal(sims.read_state(r), o.read())
    with Replicas(Opt(n_replicas=2, n_events=1000, min_loss=1e9, order="ref")) as sims:  # ZeroLoss, SimRunner.hs:161
        sims.run(1000)
        s = sims.read_summary()
        assert (s["status"] == _ffi.ZERO_LOSS).all() and (s["iter"] == 51).all()
    with Replicas(Opt(n_replicas=2, n_events=500, order="ref")) as sims:  # NaNLoss, Agent.hs:80 / SimRunner.hs:152
        sims.run(100)
        w = sims.read_state(1)["w"]
        w[:] = np.nan
        sims.write_state(1, w=w)
        sims.run(400)
        s = sims.read_summary()
        assert s["status"].tolist() == [_ffi.SUCCESS, _ffi.NAN_LOSS] and s["iter"].tolist() == [500, 101]
    with Replicas(Opt(n_events=50, order="ref", verify_reuse_constraint=True)) as sims:  # SimRunner.hs:155-156
        g = np.zeros((7, 7, 70), np.uint8)
        g[3, 3, 5] = g[3, 4, 5] = 1
        sims.write_state(0, grid=g)
        sims.run(10)
        assert sims.read_status(0)[0] == _ffi.REUSE_VIOLATED and sims.read_status(0)[1] == 1
