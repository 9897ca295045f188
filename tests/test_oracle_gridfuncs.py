"""Oracle vs test/GridfuncsSpec.hs and the src/ExGrids.hs fixture (Gridfuncs.hs)."""
import numpy as np
import pyoracle as po
import pytest


def test_inuse_channels():  # GridfuncsSpec.hs:23-30
    g = po.afterstate(po.mk_grid(), (3, 3), 13, True)
    assert po.inuse_chs(g, (3, 3)).tolist() == [13]


def test_eligible_channels():  # GridfuncsSpec.hs:32-55
    g = po.mk_grid()
    g[3, 2, 0] = 1
    assert po.eligible_chs(g, (3, 3)).tolist() == list(range(1, 70))
    g = po.afterstate(po.afterstate(po.mk_grid(), (0, 1), 69, True), (5, 0), 69, True)
    e1, e2 = po.eligible_chs(g, (0, 1)), po.eligible_chs(g, (5, 0))
    assert e1.tolist() == e2.tolist()
    assert len(e1) == 69


def test_afterstates():  # GridfuncsSpec.hs:57-65
    g = po.afterstate(po.mk_grid(), (3, 3), 0, True)
    chs = po.eligible_chs(g, (3, 3))
    afs = po.afterstates(g, (3, 3), False, chs)
    assert afs.shape == (69, 7, 7, 70)
    assert afs.sum() == 69 * 2


def _cmp_grids():
    g0 = po.mk_grid()
    g2 = po.afterstate(g0, (3, 3), 0, True)
    g3 = po.afterstate(po.afterstate(po.afterstate(g2, (3, 3), 1, True), (5, 0), 0, True), (5, 0), 1, True)
    return g0, g2, g3


def test_feature_rep_cmp():  # GridfuncsSpec.hs:72-82
    for g in _cmp_grids():
        assert np.array_equal(po.feature_rep(g), po.feature_rep_slow(g))
    f0 = po.feature_rep(po.mk_grid())
    assert f0[:, :, :70].sum() == 0 and (f0[:, :, 70] == 70).all()


def test_inc_afterstate_freps_shape():  # GridfuncsSpec.hs:84-98
    g = po.mk_grid()
    g[3, 3, 0] = 1
    chs = po.eligible_chs(g, (3, 3))
    af = po.inc_afterstate_freps((3, 3), False, chs, g, po.feature_rep(g))
    assert af.shape == (len(chs), 7, 7, 71)


def test_inc_afterstate_freps_cmp():  # GridfuncsSpec.hs:109-167 (the strongest hot-path test)
    c1, c2 = (4, 5), (6, 2)
    g0 = po.mk_grid()
    g1 = po.afterstate(g0, c1, 69, True)
    g2 = po.afterstate(g1, c2, 49, True)
    g3 = po.afterstate(g2, c2, 69, True)
    g4 = po.afterstate(g3, c2, 69, False)
    steps = [(g0, g1, c1, False, 69), (g1, g2, c2, False, 49), (g2, g3, c2, False, 69), (g3, g4, c2, True, 69)]
    for before, after, cell, is_end, ch in steps:
        fa = po.feature_rep(after)
        assert np.array_equal(fa, po.feature_rep_slow(after))
        inc = po.inc_afterstate_freps(cell, is_end, [ch], before, po.feature_rep(before))[0]
        assert np.array_equal(inc, fa)


def test_select_action_smoke():  # GridfuncsSpec.hs:169-187
    g = po.mk_grid()
    g[3, 3, 0] = 1
    f = po.feature_rep(g)
    chs = po.eligible_chs(g, (3, 3))
    w = np.zeros(po.FREP_SIZE, dtype=np.float32)
    for order in (po.ORDER_REF, po.ORDER_FAST):
        idx, qval, afrep, qvals = po.select_action(order, (3, 3), False, chs, w, g, f)
        assert idx == len(chs) - 1  # all q equal (zero weights): the LAST candidate wins
        assert qval == 0 and afrep.sum() != 0


def test_violates_reuse_constraint():  # GridfuncsSpec.hs:189-227
    g = po.mk_grid()
    assert not po.violates_reuse_constraint(g)
    g2 = po.afterstate(g, (3, 3), 3, True)
    assert not po.violates_reuse_constraint(g2)
    assert not po.violates_reuse_constraint(po.afterstate(g2, (3, 3), 4, True))
    assert po.violates_reuse_constraint(po.afterstate(g2, (3, 4), 3, True))
    assert po.violates_reuse_constraint(po.afterstate(g2, (3, 5), 3, True))
    assert not po.violates_reuse_constraint(po.afterstate(g2, (3, 6), 3, True))
    g3 = po.afterstate(po.afterstate(g, (0, 1), 69, True), (5, 0), 69, True)
    assert not po.violates_reuse_constraint(g3)


# ---- the reference's only numeric fixture: src/ExGrids.hs ------------------
def test_exgrids_frep_is_feature_rep_of_grid(exgrids):  # ExGrids.hs:25-29
    g = exgrids["prev_grid"]
    assert g.sum() == 225
    want = exgrids["prev_frep"].astype(np.int64)
    assert np.array_equal(po.feature_rep(g), want)
    assert np.array_equal(po.feature_rep_slow(g), want)
    assert not po.violates_reuse_constraint(g)


def test_exgrids_end_event_clears_the_chosen_channel(exgrids):  # ExGrids.hs:22-23,36-43
    g, cur = exgrids["prev_grid"], exgrids["cur_grid"]
    cell = tuple(int(v) for v in exgrids["prev_event_cell"])
    end_ch, pre_ch = int(exgrids["prev_event_end_ch"]), int(exgrids["pre_ch"])
    assert cell == (2, 4) and end_ch == 28 and pre_ch == 30
    inuse = po.inuse_chs(g, cell).tolist()
    assert end_ch in inuse and pre_ch in inuse and len(inuse) == 9
    # gridStep on END clears the channel the agent picked, not the ending one
    after = po.execute_action(True, cell, pre_ch, g)
    assert np.array_equal(after, cur)
    assert cur[2, 4, end_ch] == 1


def test_exgrids_incremental_equals_from_scratch_on_a_realistic_state(exgrids):
    g = exgrids["prev_grid"]
    f = exgrids["prev_frep"].astype(np.int64)
    cell = (2, 4)
    chs = po.inuse_chs(g, cell)
    inc = po.inc_afterstate_freps(cell, True, chs, g, f)
    for k, ch in enumerate(chs):
        assert np.array_equal(inc[k], po.feature_rep(po.afterstate(g, cell, int(ch), False)))
    cell = tuple(int(v) for v in exgrids["cur_event_cell"])
    g2 = exgrids["cur_grid"]
    f2 = po.feature_rep(g2)
    chs = po.eligible_chs(g2, cell)
    assert len(chs) > 0
    inc = po.inc_afterstate_freps(cell, False, chs, g2, f2)
    for k, ch in enumerate(chs):
        assert np.array_equal(inc[k], po.feature_rep(po.afterstate(g2, cell, int(ch), True)))


def test_exgrids_agent_is_consistent_with_the_dense_fill_term(exgrids):
    # Agent.hs:69: a2 = fill (c*avgR) is added to EVERY weight, so weights whose
    # feature was never non-zero (wGradCorr still exactly 0) all share one value.
    w, wg = exgrids["wnet"], exgrids["wgrad"]
    untouched = w[wg == 0.0]
    assert untouched.size > 100
    assert np.unique(untouched).size == 1
    assert untouched[0] > 0 and float(exgrids["avg_reward"]) > 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_walk_incremental_equals_from_scratch(seed):
    rng = np.random.default_rng(seed)
    g = po.mk_grid()
    f = po.feature_rep(g)
    for _ in range(300):
        cell = (int(rng.integers(7)), int(rng.integers(7)))
        inuse = po.inuse_chs(g, cell)
        is_end = bool(rng.random() < 0.45 and len(inuse))
        chs = inuse if is_end else po.eligible_chs(g, cell)
        if len(chs) == 0:
            continue
        k = int(rng.integers(len(chs)))
        f = po.inc_afterstate_freps(cell, is_end, chs, g, f)[k]
        g = po.execute_action(is_end, cell, int(chs[k]), g)
        assert not po.violates_reuse_constraint(g)
    assert np.array_equal(f, po.feature_rep(g))
    assert f.max() <= 70
