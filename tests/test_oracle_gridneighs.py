"""Oracle vs the reference's neighbourhood tests and fixtures.

Restates test/GridneighsSpec.hs:57-68 and the distance KATs of
test/GridfuncsSpec.hs:189-227; pins src/Gridneighs.hs:92-128.
"""
import pyoracle as po
import pytest

CELLS = [(r, c) for r in range(7) for c in range(7)]


def hexdist(a, b):
    dr, dc = a[0] - b[0], a[1] - b[1]
    return max(abs(dr), abs(dc), abs(dr + dc))


@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_neighborhood_with_and_without_self(d):
    # GridneighsSpec.hs:57-68: focal cell first, otherwise equal
    for cell in CELLS:
        a = po.get_neighborhood_acc(d, *cell, False)
        b = po.get_neighborhood_acc(d, *cell, True)
        assert b[0] == cell
        assert a == b[1:]


@pytest.mark.parametrize("d", [1, 2, 3, 4])
def test_table_lookup_equals_host_lists(d):
    # Gridconsts.hs:52-71 (slit of V by M) must equal Gridneighs.hs:92-98
    for cell in CELLS:
        for incl in (False, True):
            assert po.get_neighborhood_acc(d, *cell, incl) == po.get_neighs(d, *cell, incl)


def test_neighbourhood_is_axial_hex_metric():
    for cell in CELLS:
        for d in range(1, 5):
            want = {o for o in CELLS if 1 <= hexdist(cell, o) <= d}
            got = po.get_neighs(d, *cell, False)
            assert len(got) == len(set(got))
            assert set(got) == want
            ring = po.periphery(d, *cell)
            assert {hexdist(cell, o) for o in ring} <= {d}


def test_periphery_order_interior_cell():
    # Gridneighs.hs:111-121: ps6++ps5++ps4++ps3++ps2++ps1 for d=1 at (3,3)
    assert po.periphery(1, 3, 3) == [(3, 2), (4, 2), (4, 3), (3, 4), (2, 4), (2, 3)]
    # d=2: each side has 2 cells, walked from its start point
    assert po.periphery(2, 3, 3) == [
        (3, 1), (2, 2), (5, 1), (4, 1), (5, 3), (5, 2), (3, 5), (4, 4), (1, 5), (2, 5), (1, 3), (1, 4),
    ]


def test_periphery_filters_out_of_bounds():
    assert po.periphery(1, 0, 0) == [(1, 0), (0, 1)]


def test_table_sizes():
    # |_neighs| = 1539, |_neighsO| = 1490, |_neighs2| = 671
    n, no, n2, m = po.table_sizes()
    assert (n, no, n2) == (1539, 1490, 671)
    assert m[0, 0] == 0 and all(m[:, 1] >= 3)
    sizes2 = [len(po.get_neighs(2, *c, False)) for c in CELLS]
    sizes4 = [len(po.get_neighs(4, *c, False)) for c in CELLS]
    assert (min(sizes2), max(sizes2)) == (5, 18)
    assert (min(sizes4), max(sizes4)) == (14, 42)


def test_distance_kats_from_gridfuncs_spec():
    # GridfuncsSpec.hs:205-227
    assert (3, 4) in po.get_neighs(1, 3, 3, False)
    assert (3, 5) in po.get_neighs(2, 3, 3, False) and (3, 5) not in po.get_neighs(1, 3, 3, False)
    assert (3, 6) not in po.get_neighs(2, 3, 3, False) and (3, 6) in po.get_neighs(3, 3, 3, False)
    assert (5, 0) not in po.get_neighs(4, 0, 1, False)
