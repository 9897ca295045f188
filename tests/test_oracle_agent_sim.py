"""Oracle agent arithmetic (src/Agent.hs, AccUtils.hs) and whole-simulator invariants."""
import numpy as np
import pyoracle as po
import pytest

F32 = np.float32


def _state(seed=0):
    rng = np.random.default_rng(seed)
    g = po.mk_grid()
    for _ in range(150):
        cell = (int(rng.integers(7)), int(rng.integers(7)))
        chs = po.eligible_chs(g, cell)
        if len(chs):
            g[cell[0], cell[1], int(chs[rng.integers(len(chs))])] = 1
    w = rng.normal(0, 1, po.FREP_SIZE).astype(F32)
    wg = rng.normal(0, 1e-3, po.FREP_SIZE).astype(F32)
    return g, po.feature_rep(g), w, wg


def test_argpmax_last_max_wins():  # AccUtils.hs:203-208
    assert po.argpmax([1.0, 3.0, 3.0, 2.0]) == 2
    assert po.argpmax([0.0] * 7) == 6
    assert po.argpmax([5.0, 1.0]) == 0
    assert po.argpmax([4.0]) == 0


def test_dense_dot_ref_is_a_right_fold():  # Interpreter fold: x0+(x1+(...+(xn-1+0)))
    _, f, w, _ = _state(1)
    x = f.reshape(-1).astype(F32)
    acc = F32(0)
    for i in range(po.FREP_SIZE - 1, -1, -1):
        acc = F32(F32(x[i] * w[i]) + acc)
    assert po.dense_dot(po.ORDER_REF, f, w) == acc


def _fast_pieces(f, w):
    """numpy restatement of the FAST trees (oracle/dca_oracle.c dot_fast / dot_fast_g)."""
    x = f.astype(np.float64)
    w3 = w.reshape(7, 7, 71)

    def fma(a, b, c):  # exact in float64 then one rounding: |a|<=70, 24-bit b -> product exact
        return F32(np.float64(a) * np.float64(b) + np.float64(c))

    def rowdot(fe, r):  # even / odd column chains = one f32x2 accumulator pair
        e = F32(F32(x[r, 0, fe]) * w3[r, 0, fe])
        o = F32(F32(x[r, 1, fe]) * w3[r, 1, fe])
        e = fma(x[r, 2, fe], w3[r, 2, fe], e)
        o = fma(x[r, 3, fe], w3[r, 3, fe], o)
        e = fma(x[r, 4, fe], w3[r, 4, fe], e)
        o = fma(x[r, 5, fe], w3[r, 5, fe], o)
        e = fma(x[r, 6, fe], w3[r, 6, fe], e)
        return F32(e + o)

    def bfly(v):  # xor-butterfly over len(v) lanes
        v = [F32(a) for a in v]
        s = 1
        while s < len(v):
            v = [F32(v[l] + v[l ^ s]) for l in range(len(v))]
            s <<= 1
        assert all(a == v[0] or (np.isnan(a) and np.isnan(v[0])) for a in v)
        return v[0]

    def plane(fe):
        return bfly([rowdot(fe, r) for r in range(7)] + [F32(0.0)])

    return rowdot, bfly, plane


def test_dense_dot_fast_tree():
    _, f, w, _ = _state(2)
    rowdot, bfly, plane = _fast_pieces(f, w)
    v = [F32(F32(plane(l) + plane(l + 32)) + (plane(l + 64) if l < 6 else F32(0.0))) for l in range(32)]
    assert po.dense_dot(po.ORDER_FAST, f, w) == F32(bfly(v) + plane(70))


def test_dense_dot_g_fast_tree():  # x . wGradCorr: per-thread slot sums, warp butterflies, (W0 + W1) + W2
    _, f, _, wg = _state(2)
    wg = (wg + np.linspace(-1, 1, wg.size, dtype=F32)).astype(F32)
    rowdot, bfly, _ = _fast_pieces(f, wg)
    W = []
    for wp in range(3):
        v = []
        for lane in range(32):
            g, r = 4 * wp + lane // 8, lane % 8
            s = F32(0.0)
            if r < 7:
                s = rowdot(g, r)
                j = 1
                while j < 6 and 12 * j + g <= 70:
                    s = F32(s + rowdot(12 * j + g, r))
                    j += 1
            v.append(s)
        W.append(bfly(v))
    assert po.dense_dot_g(po.ORDER_FAST, f, wg) == F32(F32(W[0] + W[1]) + W[2])
    assert po.dense_dot_g(po.ORDER_REF, f, wg) == po.dense_dot(po.ORDER_REF, f, wg)


@pytest.mark.parametrize("order", [po.ORDER_REF, po.ORDER_FAST])
def test_backward_matches_numpy_restatement(order):  # Agent.hs:44-81
    g, f, w, wg = _state(3)
    cell = (3, 3)
    chs = po.eligible_chs(g, cell)
    nf = po.inc_afterstate_freps(cell, False, chs, g, f)[0]
    aN, aA, aG, avgR, reward = F32(2.52e-6), F32(0.06), F32(5e-6), F32(123.5), F32(151)
    loss, isnan, w2, wg2, avg2 = po.backward(order, aN, aA, aG, f, nf, reward, w, wg, avgR)
    val, nval, dot = po.dense_dot(order, f, w), po.dense_dot(order, nf, w), po.dense_dot_g(order, f, wg)
    td = F32(F32(F32(reward - avgR) + nval) - val)
    assert loss == td and not isnan
    c = F32(F32(-2.0) * aN)
    x, xn = f.reshape(-1).astype(F32), nf.reshape(-1).astype(F32)
    ctd, cavg, cdot = F32(c * td), F32(c * avgR), F32(c * dot)
    s = F32(aG * F32(td - dot))
    if order == po.ORDER_REF:
        grads = (ctd * x + cavg) - cdot * xn  # float32 elementwise, each op rounded
        assert grads.dtype == F32
        want_w, want_wg = w - grads, wg + s * x
    else:
        x64, xn64 = x.astype(np.float64), xn.astype(np.float64)
        t = (np.float64(ctd) * x64 + np.float64(cavg)).astype(F32)
        gr = (np.float64(-cdot) * xn64 + t.astype(np.float64)).astype(F32)
        want_w = w - gr
        want_wg = (np.float64(s) * x64 + wg.astype(np.float64)).astype(F32)
    assert np.array_equal(w2, want_w) and np.array_equal(wg2, want_wg)
    assert avg2 == F32(avgR + F32(aA * td))


def test_backward_nan_loss():  # Agent.hs:80
    g, f, w, wg = _state(4)
    w[5] = np.nan
    f = f.copy()
    f[0, 0, 5] = 1
    loss, isnan, *_ = po.backward(po.ORDER_REF, 1e-6, 0.06, 5e-6, f, f, 10.0, w, wg, 0.0)
    assert isnan and np.isnan(loss)


@pytest.mark.parametrize("order", [po.ORDER_REF, po.ORDER_FAST])
def test_get_action_and_grid_step_train(order):  # Simulator.hs:154-172, SimRunner.hs:100-118
    g, f, w, wg = _state(5)
    cell = (2, 4)
    ch, nf, ncand = po.get_action(order, cell, False, w, g, f)
    chs = po.eligible_chs(g, cell)
    assert ncand == len(chs) and ch in chs
    idx, qval, afrep, qvals = po.select_action(order, cell, False, chs, w, g, f)
    assert chs[idx] == ch and np.array_equal(afrep, nf)
    assert qval == qvals.max() and idx == max(i for i, q in enumerate(qvals) if q == qval)
    g2, w2, wg2, avg2, loss, isnan, reward = po.grid_step_train(order, 2.52e-6, 0.06, 5e-6, False, cell, ch, f, nf, g, w, wg, 0.0)
    assert g2[cell[0], cell[1], ch] == 1 and reward == g.sum() + 1
    assert np.array_equal(nf, po.feature_rep(g2))
    # END: candidates are the in-use channels; the chosen one is cleared
    ch_e, nf_e, ncand_e = po.get_action(order, cell, True, w2, g2, nf)
    assert ncand_e == len(po.inuse_chs(g2, cell)) and g2[cell[0], cell[1], ch_e] == 1
    g3, *_rest, reward3 = po.grid_step_train(order, 2.52e-6, 0.06, 5e-6, True, cell, ch_e, nf, nf_e, g2, w2, wg2, avg2)
    assert g3[cell[0], cell[1], ch_e] == 0 and reward3 == reward - 1
    # no candidates: (Nothing, frep) and the grid is untouched
    full = po.mk_grid()
    full[3, 3, :] = 1
    ch_n, nf_n, nc = po.get_action(order, (3, 4), False, w, full, po.feature_rep(full))
    assert (ch_n, nc) == (-1, 0) and np.array_equal(nf_n, po.feature_rep(full))


@pytest.mark.parametrize("order", [po.ORDER_REF, po.ORDER_FAST])
@pytest.mark.parametrize("hoff", [0.0, 0.15])
def test_simulation_invariants(order, hoff):
    n = 3000
    s = po.Sim(order=order, hoff_prob=hoff, n_events=n)
    status, done, tr = s.run(n, trace=True)
    st = s.read()
    assert status == po.SUCCESS and done == n and st["iter"] == n
    arrN, accN, rejN, arrH, ended, rejH = st["stats"][:6]
    assert rejH == 0  # statsEventRejectHoff is never called (Simulator.hs:113)
    # Stats.hs:139-147 call conservation
    assert arrN + arrH - rejN - rejH - ended == st["grid"].sum() == tr["reward"][-1]
    assert not po.violates_reuse_constraint(st["grid"])
    assert np.array_equal(st["frep"], po.feature_rep(st["grid"]))  # incremental == from scratch
    assert st["frep"].max() <= 70
    assert (np.diff(tr["time"]) >= 0).all()
    assert int((tr["etype"] == po.NEW).sum()) == arrN and int((tr["etype"] == po.END).sum()) == ended
    assert tr["grid_hash"][-1] == po.grid_hash(st["grid"]) and tr["frep_hash"][-1] == po.frep_hash(st["frep"])
    if hoff == 0:
        assert arrH == 0 and accN - ended == st["grid"].sum()
    else:
        assert arrH > 50
        # every END with a hand-off target is immediately followed by its HOFF at the same time
        hoffs = np.flatnonzero(tr["etype"] == po.HOFF)
        assert (tr["etype"][hoffs - 1] == po.END).all() and (tr["time"][hoffs - 1] == tr["time"][hoffs]).all()


def test_config1_sanity_ranges():
    # BASELINE.md section 4 (survey scratch restatement; sanity ranges, not parity targets)
    s = po.Sim(order=po.ORDER_REF)
    status, done, tr = s.run(10000, trace=True)
    st = s.read()
    assert status == po.SUCCESS
    assert st["stats"][0] == pytest.approx(5500, abs=60)
    assert 500 <= st["stats"][2] <= 660
    assert 420 <= st["grid"].sum() <= 450
    assert 15300 <= s.ndraws <= 15500
    assert 8.3 <= tr["ncand"].mean() <= 9.0
    assert 19.5 < st["w"].min() and st["w"].max() < 23.5
    assert po.stats_cumus(st["stats"])[0] == pytest.approx(st["stats"][2] / (st["stats"][0] + 1.0))


def test_period_loss_accumulator_and_stop_rules():
    s = po.Sim(order=po.ORDER_REF, n_events=2500, log_iter=1000)
    status, done, tr = s.run(1000, trace=True)
    assert status == po.RUNNING and done == 1000
    lsum, ln = s.period(reset=True)
    acc = np.float32(0)
    for v in tr["loss"]:
        acc = np.float32(acc + v)
    assert ln == 1000 and lsum == acc
    assert s.read()["stats"][6] == 0  # period counters reset (Stats.hs:117-118)
    status, done, _ = s.run(5000)
    assert status == po.SUCCESS and done == 1500
    status, done, _ = s.run(10)
    assert done == 0
    z = po.Sim(order=po.ORDER_REF, min_loss=1e9)
    status, done, _ = z.run(1000)
    assert status == po.ZERO_LOSS and done == 51  # iTotal > 50 (SimRunner.hs:161)


def test_run_replicas_philox_is_deterministic_and_distinct():
    o = po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=3)
    t1, st1, avg1, bp1 = po.run_replicas(o, 0, 4, 400, 2)
    t2, st2, avg2, bp2 = po.run_replicas(o, 0, 4, 400, 1)
    assert t1 == t2 == 1600 and np.array_equal(st1, st2) and np.array_equal(avg1, avg2)
    assert len(set(avg1.tolist())) == 4
    o.replica = 2
    s = po.Sim(o)
    s.opt.n_events = 400
    s2 = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=3, replica=2, n_events=400))
    s2.run(400)
    assert s2.read()["avg_reward"] == avg1[2]
