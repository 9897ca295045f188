"""Short profiling target for ncu: the fused FAST kernel on a batch of replicas.
usage: prof_run.py [n_replicas] [n_events] [warm_events]
  warm_events = 0: two launches of n_events from the empty grid (includes the 70-candidate transient)
  warm_events > 0: one launch of warm_events, then one of n_events from the warmed-up state
                   (profile the second one: ncu -k regex:k_run_fast -s 1 -c 1)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haskelldca_b200 import Opt, Replicas

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with Replicas(Opt(n_replicas=n_rep, n_events=warm + n_ev if warm else n_ev, order="fast", rng="philox")) as s:
    if warm:
        s.run(warm)
        ms = s.run(n_ev)
        total = n_rep * (warm + n_ev)
    else:
        for _ in range(2):
            s.reset()
            ms = s.run(n_ev)
        total = n_rep * n_ev
    tot = s.sum_stats()[0]
    assert tot[8] == total and tot[9] == 0, tot
    print(f"{n_rep} x {n_ev}{' (warm %d)' % warm if warm else ''}: {ms:.2f} ms, {n_rep*n_ev/ms/1e3:.2f} M events/s")
