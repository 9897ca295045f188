"""Tiny run for compute-sanitizer (memcheck / racecheck): FAST kernel with hand-offs, plain and traced."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haskelldca_b200 import Opt, Replicas

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for trace in (0, n_ev):
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", hoff_prob=0.15, trace=trace,
                      verify_reuse_constraint=bool(trace))) as s:
        s.run(n_ev // 2)
        s.run(n_ev)
        tot = s.sum_stats()[0]
        assert tot[8] == n_rep * n_ev and tot[9] == 0, tot
print("ok")
