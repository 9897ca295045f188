"""Device Gridfuncs/Gridneighs operators vs test/GridfuncsSpec.hs KATs, src/ExGrids.hs and the oracle."""
import numpy as np
import pytest

import pyoracle as po

pytestmark = pytest.mark.gpu


def test_inuse_and_eligible_kats():  # GridfuncsSpec.hs:23-55
    from haskelldca_b200 import gridfuncs as gf

    g = gf.mk_grid()
    g[3, 3, 13] = 1
    assert gf.inuse_chs((3, 3), g).tolist() == [13]
    g = gf.mk_grid()
    g[3, 2, 0] = 1
    assert gf.eligible_chs((3, 3), g).tolist() == list(range(1, 70))
    g = gf.mk_grid()
    g[0, 1, 69] = g[5, 0, 69] = 1
    e1, e2 = gf.eligible_chs((0, 1), g), gf.eligible_chs((5, 0), g)
    assert e1.tolist() == e2.tolist() and len(e1) == 69


def test_feature_rep_and_incremental_cmp():  # GridfuncsSpec.hs:72-82,109-167
    from haskelldca_b200 import gridfuncs as gf

    c1, c2 = (4, 5), (6, 2)
    g0 = gf.mk_grid()
    g1 = po.afterstate(g0, c1, 69, True)
    g2 = po.afterstate(g1, c2, 49, True)
    g3 = po.afterstate(g2, c2, 69, True)
    g4 = po.afterstate(g3, c2, 69, False)
    for before, after, cell, is_end, ch in [(g0, g1, c1, False, 69), (g1, g2, c2, False, 49), (g2, g3, c2, False, 69), (g3, g4, c2, True, 69)]:
        fa = gf.feature_rep(after)
        assert np.array_equal(fa, po.feature_rep(after)) and np.array_equal(fa, po.feature_rep_slow(after))
        inc = gf.inc_afterstate_freps(cell, is_end, [ch], before, gf.feature_rep(before))
        assert inc.shape == (1, 7, 7, 71) and np.array_equal(inc[0], fa)


def test_violates_reuse_constraint_kats():  # GridfuncsSpec.hs:189-227
    from haskelldca_b200 import gridfuncs as gf

    g = gf.mk_grid()
    assert not gf.violates_reuse_constraint(g)
    g2 = po.afterstate(g, (3, 3), 3, True)
    assert not gf.violates_reuse_constraint(g2)
    assert not gf.violates_reuse_constraint(po.afterstate(g2, (3, 3), 4, True))
    assert gf.violates_reuse_constraint(po.afterstate(g2, (3, 4), 3, True))
    assert gf.violates_reuse_constraint(po.afterstate(g2, (3, 5), 3, True))
    assert not gf.violates_reuse_constraint(po.afterstate(g2, (3, 6), 3, True))


def test_exgrids_fixture_on_device(exgrids):  # src/ExGrids.hs:22-43
    from haskelldca_b200 import gridfuncs as gf

    g, f = exgrids["prev_grid"], exgrids["prev_frep"].astype(np.int64)
    assert np.array_equal(gf.feature_rep(g), f)
    chs = gf.inuse_chs((2, 4), g)
    assert chs.tolist() == po.inuse_chs(g, (2, 4)).tolist() and 28 in chs and 30 in chs
    inc = gf.inc_afterstate_freps((2, 4), True, chs, g, f)
    assert np.array_equal(inc, po.inc_afterstate_freps((2, 4), True, chs, g, f))
    cur = exgrids["cur_grid"]
    chs = gf.eligible_chs((3, 3), cur)
    assert chs.tolist() == po.eligible_chs(cur, (3, 3)).tolist()
    fc = gf.feature_rep(cur)
    assert np.array_equal(gf.inc_afterstate_freps((3, 3), False, chs, cur, fc), po.inc_afterstate_freps((3, 3), False, chs, cur, fc))


def test_random_states_against_the_literal_nested_scan():
    """The device's cnt2 formulation of Gridfuncs.hs:220-232 vs the oracle's literal neighbour scan."""
    from haskelldca_b200 import gridfuncs as gf

    rng = np.random.default_rng(5)
    g = gf.mk_grid()
    for step in range(120):
        cell = (int(rng.integers(7)), int(rng.integers(7)))
        inuse = po.inuse_chs(g, cell)
        is_end = bool(rng.random() < 0.4 and len(inuse))
        chs = inuse if is_end else po.eligible_chs(g, cell)
        if len(chs) == 0:
            continue
        if step % 6 == 0:
            f = po.feature_rep(g)
            dev_chs = gf.inuse_chs(cell, g) if is_end else gf.eligible_chs(cell, g)
            assert dev_chs.tolist() == chs.tolist()
            assert np.array_equal(gf.inc_afterstate_freps(cell, is_end, chs, g, f), po.inc_afterstate_freps(cell, is_end, chs, g, f))
            assert np.array_equal(gf.feature_rep(g), f)
        g = po.execute_action(is_end, cell, int(chs[rng.integers(len(chs))]), g)


def test_neighbourhood_order_matches_reference_walk():
    from haskelldca_b200 import gridneighs

    for cell in [(r, c) for r in range(7) for c in range(7)]:
        for d in (1, 2, 3, 4):
            for incl in (False, True):
                assert gridneighs.get_neighs(d, cell, incl) == po.get_neighs(d, *cell, incl)
