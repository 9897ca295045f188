"""Hand-off look-ahead (extension; the reference's only TODO, README.md:71-72) in the oracle: definition checks on
the CPU.  q[k] = max over h eligible at the target cell on afterstate_k of V(afterstate_k + h), V(afterstate_k) when
nothing is eligible; argpmax over k; the returned frep is the plain afterstate of the chosen channel."""
import numpy as np
import pytest

import pyoracle as po

ORDERS = [po.ORDER_REF, po.ORDER_FAST]


def _state(seed, n_calls=260):
    rng = np.random.default_rng(seed)
    g = po.mk_grid()
    placed = 0
    for _ in range(40 * n_calls):
        r, c = int(rng.integers(7)), int(rng.integers(7))
        chs = po.eligible_chs(g, (r, c))
        if len(chs):
            g[r, c, int(rng.choice(chs))] = 1
            placed += 1
            if placed == n_calls:
                break
    return g, po.feature_rep(g), rng.normal(20.0, 3.0, po.FREP_SIZE).astype(np.float32), rng


@pytest.mark.parametrize("order", ORDERS)
def test_lookahead_values_from_first_principles(order):
    g, f, w, rng = _state(3)
    done = 0
    for trial in range(200):
        cell = (int(rng.integers(7)), int(rng.integers(7)))
        inuse = po.inuse_chs(g, cell)
        if len(inuse) < 2:
            continue
        neighs = po.get_neighs(2, cell[0], cell[1], False)
        to = neighs[int(rng.integers(len(neighs)))]
        ch, nf, ncand, q = po.get_action_lookahead(order, cell, to, w, g, f)
        assert ncand == len(inuse)
        want_q = []
        for k in inuse:
            gk = po.afterstate(g, cell, int(k), False)
            fk = po.feature_rep(gk)  # == the incremental afterstate frep (GridfuncsSpec.hs:109-167)
            hs = po.eligible_chs(gk, to)
            if len(hs) == 0:
                want_q.append(po.dense_dot(order, fk, w))
            else:
                vals = [po.dense_dot(order, po.feature_rep(po.afterstate(gk, to, int(h), True)), w) for h in hs]
                want_q.append(max(vals))
        want_q = np.array(want_q, dtype=np.float32)
        assert np.array_equal(q.view(np.uint32), want_q.view(np.uint32)), (trial, cell, to)
        idx = po.argpmax(want_q)
        assert ch == inuse[idx]
        assert np.array_equal(nf, po.feature_rep(po.afterstate(g, cell, int(ch), False)))
        done += 1
        if done == 6:
            break
    assert done == 6


def test_lookahead_changes_only_handoff_departures():
    """without hand-offs the flag changes nothing; with hand-offs the runs diverge, the invariants hold"""
    n = 3000
    a = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=4, n_events=n, hoff_lookahead=0))
    b = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=4, n_events=n, hoff_lookahead=1))
    _, _, ta = a.run(n, trace=True)
    _, _, tb = b.run(n, trace=True)
    assert np.array_equal(ta["ch"], tb["ch"]) and np.array_equal(ta["grid_hash"], tb["grid_hash"])
    a = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=4, n_events=n, hoff_prob=0.3, hoff_lookahead=0))
    b = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=4, n_events=n, hoff_prob=0.3, hoff_lookahead=1))
    _, _, ta = a.run(n, trace=True)
    sb, db, tb = b.run(n, trace=True)
    assert (sb, db) == (po.SUCCESS, n) and not np.array_equal(ta["ch"], tb["ch"])
    rb = b.read()
    st = rb["stats"]
    assert not po.violates_reuse_constraint(rb["grid"])
    assert rb["grid"].sum() == st[0] + st[3] - st[2] - st[5] - st[4]  # call conservation, Stats.hs:139
    assert np.array_equal(rb["frep"], po.feature_rep(rb["grid"]))
