"""Batched throughput of the reference-order run kernel (dca::k_run<ORDER_REF>) from a warmed-up state.
usage: ref_run.py [n_replicas] [n_events] [warm_events] [order]
The warm-up runs in the same order (the 70-candidate transient of the empty grid is over after ~3000 events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haskelldca_b200 import Opt, Replicas

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
order = sys.argv[4] if len(sys.argv) > 4 else "ref"
with Replicas(Opt(n_replicas=n_rep, n_events=warm + n_ev, order=order, rng="philox")) as s:
    ms_w = s.run(warm) if warm else 0.0
    ms = s.run(n_ev)
    tot = s.sum_stats()[0]
    assert tot[8] == n_rep * (warm + n_ev) and tot[9] == 0, tot
    print(f"{order} {n_rep} x {n_ev} (warm {warm}: {ms_w:.1f} ms): {ms:.2f} ms, {n_rep*n_ev/ms/1e3:.2f} M events/s, "
          f"{ms*1e3/n_ev:.2f} us per event step")
