"""Short profiling target for ncu: a few launches of the fused FAST kernel on a small batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haskelldca_b200 import Opt, Replicas

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox")) as s:
    for _ in range(2):
        s.reset()
        ms = s.run(n_ev)
    tot = s.sum_stats()[0]
    assert tot[8] == n_rep * n_ev and tot[9] == 0
    print(f"{n_rep} x {n_ev}: {ms:.2f} ms, {n_rep*n_ev/ms/1e3:.2f} M events/s")
