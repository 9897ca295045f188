"""Throughput sweep helper (development tool): events/s of the fused kernel for a few CTA widths."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from haskelldca_b200 import Opt, Replicas

def run(n_rep, n_ev, warps, hoff=0.0, reps=2):
    with Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", warps_per_replica=warps, hoff_prob=hoff)) as s:
        best = 1e30
        for _ in range(reps):
            s.reset()
            ms = s.run(n_ev)
            best = min(best, ms)
        tot = s.sum_stats()[0]
        assert tot[8] == n_rep * n_ev and tot[9] == 0
        return n_rep * n_ev / (best / 1e3), best

if __name__ == "__main__":
    n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    for w in (1, 2, 4):
        v, ms = run(n_rep, n_ev, w)
        print(f"replicas={n_rep} events={n_ev} warps={w}: {v/1e6:.2f} M events/s  ({ms:.1f} ms)  roofline_frac={v*55664/6453.1e9:.3f}", flush=True)
