"""Soak of the REF-order run kernel: BASELINE configs[2] (4096 replicas x 100 000 events, Philox seed 0) in the
reference's arithmetic, hand-offs on, in periods like runPeriod; at the end every replica must have reached Success,
conserve its calls (arrivals - rejected - ended == calls in progress, Stats.hs:139) and a sample of replicas must equal
their CPU oracle runs bit for bit.  usage: ref_soak.py [n_replicas] [n_events] [hoff_prob]"""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import pyoracle as po
from haskelldca_b200 import Opt, Replicas, _ffi

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
hoff = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
spot = sorted({0, 443, 444, n_rep // 2, n_rep - 1})
with ThreadPoolExecutor(max_workers=len(spot)) as ex:
    def one(r):
        o = po.Sim(po.default_opt(order=po.ORDER_REF, rng_mode=po.RNG_PHILOX, seed=0, replica=r, n_events=n_ev, hoff_prob=hoff))
        o.run(n_ev)
        return o.read()
    futs = [ex.submit(one, r) for r in spot]  # the oracle runs on host threads while the GPU works
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="ref", rng="philox", rng_seed=0, hoff_prob=hoff)) as s:
        ms = 0.0
        done = 0
        while done < n_ev:
            k = min(25000, n_ev - done)
            ms += s.run(k)
            done += k
        summ = s.read_summary()
        assert (summ["status"] == _ffi.SUCCESS).all() and (summ["iter"] == n_ev).all(), "not every replica reached Success"
        st = summ["stats"].astype(np.int64)
        tot = s.sum_stats()[0]
        assert tot[8] == n_rep * n_ev and tot[9] == 0
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        for r, f in zip(spot, futs):
            want, got = f.result(), s.read_state(r)
            for key in ("grid", "frep", "stats"):
                assert np.array_equal(got[key], want[key]), (r, key)
            for key in ("w", "wg", "avg_reward"):
                assert np.array_equal(u32(got[key]), u32(want[key])), (r, key)
            calls = int(np.asarray(got["grid"]).sum())
            sr = got["stats"]
            assert int(sr[0]) + int(sr[3]) - int(sr[2]) - int(sr[4]) == calls, (r, "call conservation", sr, calls)
        bp = st[:, 2].sum() / (st[:, 0].sum() + 1.0)
        print(f"ref soak ok: {n_rep} x {n_ev} events (hoff_prob {hoff}) in {ms:.0f} ms = {n_rep * n_ev / ms / 1e3:.2f} M events/s; "
              f"replicas {spot} bit-exact vs the oracle; cumulative blocking probability {bp:.4f}")
