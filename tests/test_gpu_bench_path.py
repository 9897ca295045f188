"""The BENCHMARKED instantiation against the oracle: BASELINE configs[2] exactly as bench.py runs it
(4096 replicas, Philox seed 0, hoff_prob 0, no trace, no verify flag => dca::fast::k_run_fast<false>).

The event role of a CTA rotates over the hardware warps from wave to wave (role_rotation(),
csrc/dca_fast.cuh: blockIdx / (5 x #SM) mod 4), so one replica of EVERY wave / rotation is compared
bit for bit with its own oracle run (runStep, src/SimRunner.hs:60-97): grid, frep, statistics, wNet,
wGradCorr, average reward, next event, number of random draws.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import pyoracle as po

pytestmark = pytest.mark.gpu

N_REPLICAS = 4096
# rotation 0 (first wave: 0..739 on a 148-SM B200), 1 (740..1479), 2, 3, then 0 again, and the last block
SPOT = (0, 739, 740, 1480, 2220, 2960, 3700, 4095)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _oracle(replica, n_total, **kw):
    return po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=0, replica=replica, n_events=n_total, **kw))


def _assert_same(sims, r, o, first=0):
    got, want = sims.read_state(r), o.read()
    for k in ("grid", "frep", "stats"):
        assert np.array_equal(got[k], want[k]), (first + r, k)
    assert got["iter"] == want["iter"] and got["event"] == want["event"], first + r
    for k in ("w", "wg", "avg_reward"):
        assert np.array_equal(_bits(got[k]), _bits(want[k])), (first + r, k)
    assert sims.read_status(r)[2] == o.ndraws, first + r


def _advance(oracles, n):
    """ctypes releases the GIL: the oracle runs of the spot replicas proceed on separate host threads"""
    with ThreadPoolExecutor(max_workers=len(oracles)) as ex:
        list(ex.map(lambda o: o.run(n), oracles.values()))


def test_bench_configuration_matches_the_oracle_in_every_wave():
    from haskelldca_b200 import Opt, Replicas

    n_total = 100000
    with Replicas(Opt(n_replicas=N_REPLICAS, n_events=n_total, order="fast", rng="philox", rng_seed=0)) as sims:
        assert sims.opt.trace == 0 and not sims.opt.verify_reuse_constraint  # => k_run_fast<false>
        oracles = {r: _oracle(r, n_total) for r in SPOT}
        sims.run(2000)
        _advance(oracles, 2000)
        for r in SPOT:
            _assert_same(sims, r, oracles[r])
        # ... and the full 100 000 events of the bench step for one replica of each rotation
        long_run = (0, 740, 2220, 4095)
        sims.run(n_total - 2000)
        summ = sims.read_summary()
        assert (summ["status"] == po.SUCCESS).all() and (summ["iter"] == n_total).all()
        _advance({r: oracles[r] for r in long_run}, n_total - 2000)
        for r in long_run:
            _assert_same(sims, r, oracles[r])


def test_bench_configuration_with_a_shard_offset_and_handoffs():
    """What rank k of a multi-GPU bench runs: the same kernel on replicas [first, first + n) -- the Philox stream
    is the GLOBAL replica index -- here with hand-offs enabled so that every event kind is on the path."""
    from haskelldca_b200 import Opt, Replicas

    first, n_rep, n_ev = 3 * 4096, 2400, 3000
    spot = (0, 739, 740, 1480, 2220, 2399)
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=0, hoff_prob=0.15,
                      first_replica=first)) as sims:
        sims.run(n_ev)
        oracles = {r: _oracle(first + r, n_ev, hoff_prob=0.15) for r in spot}
        _advance(oracles, n_ev)
        for r in spot:
            _assert_same(sims, r, oracles[r], first)
