"""Oracle EventGen / RNG vs test/EventGenSpec.hs and published known answers."""
import math

import numpy as np
import pyoracle as po
import pytest


# ---- RNG -------------------------------------------------------------------
def test_mt19937_64_known_answers():
    # Matsumoto & Nishimura mt19937-64; C++ [rand.predef]: the 10000th output of
    # the default-seeded (5489) mt19937_64 is 9981545732273789042.
    r = po.Rng(po.RNG_MT, 5489)
    first = r.word64()
    for _ in range(9998):
        r.word64()
    assert first == 14514284786278117030
    assert r.word64() == 9981545732273789042
    r0 = po.Rng(po.RNG_MT, 0)  # --fixed_rng => seed 0 (Opt.hs:79-82)
    assert [r0.word64(), r0.word64()] == [2947667278772165694, 18301848765998365067]


def test_random_double_is_53_bit():
    r, r2 = po.Rng(po.RNG_MT, 0), po.Rng(po.RNG_MT, 0)
    for _ in range(100):
        w = r.word64()
        assert r2.double() == (w >> 11) / 2.0 ** 53


def test_exponential_is_neglog_times_mean():
    r, r2 = po.Rng(po.RNG_MT, 7), po.Rng(po.RNG_MT, 7)
    for _ in range(100):
        u = r.double()
        assert r2.exponential(0.3) == (-math.log(u)) * 0.3


def test_uniform_int_rejection_on_low_byte():
    r, r2 = po.Rng(po.RNG_MT, 3), po.Rng(po.RNG_MT, 3)
    for m in (5, 7, 12, 18):
        for _ in range(200):
            got = r.uniform_int(0, m - 1)
            while True:
                z = r2.word64() & 0xFF
                if z >= 256 % m:
                    break
            assert got == z % m
    assert r.ndraws == r2.ndraws


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors (philox4x32 10 rounds)
    assert po.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert po.philox4x32_10([f, f, f, f], [f, f]).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert po.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]).tolist() == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_stream_is_counter_based():
    a, b = po.Rng(po.RNG_PHILOX, 11, 5), po.Rng(po.RNG_PHILOX, 11, 6)
    wa = [a.word64() for _ in range(4)]
    wb = [b.word64() for _ in range(4)]
    assert wa != wb
    out = po.philox4x32_10([2, 0, 5, 0], [11, 0])
    assert wa[2] == int(out[0]) | (int(out[1]) << 32)


def test_neglog_spec_matches_libm_closely():
    rng = np.random.default_rng(0)
    words = [int(x) for x in rng.integers(0, 2 ** 63, size=2000, dtype=np.uint64)] + [1 << 11, (1 << 64) - 1, 3 << 40]
    for w in words:
        u = (w >> 11) / 2.0 ** 53
        got = po.neglog_u53(w)
        want = -math.log(u)
        assert got == pytest.approx(want, rel=1e-14, abs=1e-16)
        assert got >= 0
    assert po.neglog_u53(0) == math.inf and po.neglog_u53(2047) == math.inf


# ---- EventGen (test/EventGenSpec.hs) ---------------------------------------
def test_push_new_event_to_empty_eventgen():  # EventGenSpec.hs:30-38
    eg = po.EventGen()
    eg.push(1.3, po.NEW, (1, 2))
    assert eg.eg_id == 1
    assert eg.sizes() == (1, 1, 0)


def test_push_then_pop_returns_same_event():  # :54-57
    eg = po.EventGen()
    eg.push(1.3, po.NEW, (1, 2))
    assert eg.pop().astuple() == (1.3, po.NEW, 1, 2, -1, -1, -1)


def test_generate_new_then_pop():  # :59-65
    eg = po.EventGen()
    eg.generate_new(1.3, (1, 2))
    ev = eg.pop()
    assert (ev.r, ev.c, ev.etype) == (1, 2, po.NEW)
    assert ev.time > 1.3
    assert eg.is_empty()


def test_fresh_generator_has_exactly_49_events():  # :77-91
    eg = po.EventGen(with_initial_events=True)
    assert eg.eg_id == 49 and eg.ndraws == 49
    evs = [eg.pop() for _ in range(49)]
    assert sorted((e.r, e.c) for e in evs) == [(r, c) for r in range(7) for c in range(7)]
    times = [e.time for e in evs]
    assert times == sorted(times)
    with pytest.raises(IndexError):
        eg.pop()


def test_initial_events_follow_the_seed0_stream():
    # mkEventGen, EventGen.hs:54-58: one NEW per cell in row-major order at 0 + Exp(60/rate)
    eg = po.EventGen(with_initial_events=True)
    r = po.Rng(po.RNG_MT, 0)
    want = {(i // 7, i % 7): (-math.log(r.double())) * (1.0 / (200.0 / 60.0)) for i in range(49)}
    for _ in range(49):
        e = eg.pop()
        assert e.time == 0.0 + want[(e.r, e.c)]


def test_reassign_changes_only_the_end_channel():  # :93-108
    eg = po.EventGen()
    eg.push(1.3, po.END, (1, 2), ch=7)
    eg.reassign((1, 2), 7, 13)
    ev = eg.pop()
    assert ev.astuple() == (1.3, po.END, 1, 2, 13, -1, -1)
    assert eg.is_empty()
    eg.push(2.0, po.END, (1, 2), ch=7, hoff=(1, 3))
    eg.reassign((1, 2), 7, 9)
    assert eg.pop().astuple() == (2.0, po.END, 1, 2, 9, 1, 3)  # hoffCell preserved (EventGen.hs:71)
    with pytest.raises(KeyError):
        eg.reassign((1, 2), 7, 13)


def test_equal_times_pop_in_id_order():  # EventGen/Internal.hs:29-34
    eg = po.EventGen()
    eg.push(5.0, po.HOFF, (0, 0))
    eg.push(5.0, po.NEW, (1, 1))
    eg.push(4.0, po.NEW, (2, 2))
    assert [(e.r, e.c) for e in (eg.pop(), eg.pop(), eg.pop())] == [(2, 2), (0, 0), (1, 1)]


def test_hoff_new_pushes_end_then_hoff_at_same_time():  # EventGen.hs:94-109
    eg = po.EventGen(po.default_opt(seed=4))
    eg.generate_hoff_new(1.0, (3, 3), 17)
    a, b = eg.pop(), eg.pop()
    assert a.etype == po.END and b.etype == po.HOFF and a.time == b.time > 1.0
    assert (a.r, a.c, a.ch) == (3, 3, 17)
    assert (a.hoff_r, a.hoff_c) == (b.r, b.c)
    assert (b.r, b.c) in po.get_neighs(2, 3, 3, False)
    assert eg.eg_id == 2 and eg.ndraws >= 2
