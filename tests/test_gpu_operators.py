"""accFn1 / accFn2 on explicit arrays (the pure operators runStep receives, src/SimRunner.hs:61-62)
against the oracle, both arithmetic orders, bit-exact."""
import numpy as np
import pytest

import pyoracle as po
from haskelldca_b200 import agent, gridfuncs as gf

pytestmark = pytest.mark.gpu
F32 = np.float32
ORDERS = [("ref", po.ORDER_REF), ("fast", po.ORDER_FAST)]


def _random_state(seed, n_calls=300):
    """a reachable grid (reuse constraint holds), its frep and random weights"""
    rng = np.random.default_rng(seed)
    g = po.mk_grid()
    placed = 0
    for _ in range(20 * n_calls):
        r, c = int(rng.integers(7)), int(rng.integers(7))
        chs = po.eligible_chs(g, (r, c))
        if len(chs):
            g[r, c, int(rng.choice(chs))] = 1
            placed += 1
            if placed == n_calls:
                break
    f = po.feature_rep(g)
    w = rng.normal(20.0, 3.0, po.FREP_SIZE).astype(F32)
    wg = rng.normal(0.0, 1e-3, po.FREP_SIZE).astype(F32)
    return g, f, w, wg, rng


@pytest.mark.parametrize("order,oorder", ORDERS)
def test_dense_dots(order, oorder):  # vvMul / forward, Agent.hs:22-23,62-66
    g, f, w, wg, _ = _random_state(1)
    assert agent.forward(f, w, order).view(np.uint32) == po.dense_dot(oorder, f, w).view(np.uint32)
    assert agent.dot_grad_corr(f, wg, order).view(np.uint32) == po.dense_dot_g(oorder, f, wg).view(np.uint32)


@pytest.mark.parametrize("order,oorder", ORDERS)
def test_get_action_matches_oracle(order, oorder):  # Simulator.hs:154-172
    g, f, w, wg, rng = _random_state(2)
    seen_empty = False
    for trial in range(12):
        cell = (int(rng.integers(7)), int(rng.integers(7)))
        is_end = bool(trial % 2)
        ch, nf = agent.get_action(cell, is_end, w, g, f, order)
        och, onf, ncand = po.get_action(oorder, cell, is_end, w, g, f)
        assert ch == och and np.array_equal(nf, onf), (trial, cell, is_end)
        seen_empty |= ncand == 0
    # no candidate: (Nothing, frep)
    full = po.mk_grid()
    full[3, 3, :] = 1
    ff = po.feature_rep(full)
    ch, nf = agent.get_action((3, 4), False, w, full, ff, order)
    assert ch == -1 and np.array_equal(nf, ff)
    ch, nf = agent.get_action((0, 0), True, w, full, ff, order)
    assert ch == -1 and np.array_equal(nf, ff)


@pytest.mark.parametrize("order,oorder", ORDERS)
def test_frep_argument_is_honoured(order, oorder):
    """accFn1 takes frep as an argument (it need not equal featureRep grid): Gridfuncs.hs:206 replicates it"""
    g, f, w, wg, rng = _random_state(3)
    f2 = f + rng.integers(0, 3, f.shape)
    ch, nf = agent.get_action((2, 5), False, w, g, f2, order)
    och, onf, _ = po.get_action(oorder, (2, 5), False, w, g, f2)
    assert ch == och and np.array_equal(nf, onf)


@pytest.mark.parametrize("order,oorder", ORDERS)
def test_grid_step_train_matches_oracle(order, oorder):  # SimRunner.hs:100-118, Agent.hs:44-81
    g, f, w, wg, rng = _random_state(4)
    avg = F32(412.25)
    aN, aA, aG = 2.52e-6, 0.06, 5e-6
    for trial in range(6):
        cell = (int(rng.integers(7)), int(rng.integers(7)))
        is_end = bool(trial % 2)
        ch, nf = agent.get_action(cell, is_end, w, g, f, order)
        g2, (w2, wg2, avg2), loss, reward = agent.grid_step_train(aN, aA, aG, is_end, cell, ch if ch >= 0 else None, f, nf, g, (w, wg, avg), order)
        og, ow, owg, oavg, oloss, onan, oreward = po.grid_step_train(oorder, aN, aA, aG, is_end, cell, ch, f, nf, g, w, wg, avg)
        assert np.array_equal(g2, og) and reward == oreward
        assert np.array_equal(w2.view(np.uint32), ow.view(np.uint32)) and np.array_equal(wg2.view(np.uint32), owg.view(np.uint32))
        assert F32(avg2).view(np.uint32) == F32(oavg).view(np.uint32) and (loss is None) == bool(onan)
        if loss is not None:
            assert F32(loss).view(np.uint32) == F32(oloss).view(np.uint32)
        g, f, w, wg, avg = g2, nf, w2, wg2, avg2
        assert np.array_equal(f, gf.feature_rep(g))  # the chosen afterstate frep is featureRep of the new grid


def test_nan_loss_is_nothing():  # Agent.hs:80
    g, f, w, wg, _ = _random_state(5, n_calls=50)
    w = w.copy()
    w[5] = np.nan
    f = f.copy()
    f[0, 0, 5] = 1
    _, _, loss, _ = agent.grid_step_train(1e-6, 0.06, 5e-6, False, (0, 0), None, f, f, g, (w, wg, 0.0), "ref")
    assert loss is None


def test_bad_arguments_are_errors():
    from haskelldca_b200._ffi import DcaError
    g, f, w, wg, _ = _random_state(6, n_calls=10)
    with pytest.raises(DcaError):
        agent.get_action((7, 0), False, w, g, f)
    with pytest.raises(ValueError):
        agent.get_action((0, 0), False, w[:10], g, f)
