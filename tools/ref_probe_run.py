"""Timing only (no checks): the REF-order run kernel of the library in DCA_LIB on a warmed-up state.
usage: ref_probe_run.py n_replicas n_events warm_events"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haskelldca_b200 import Opt, Replicas
n_rep, n_ev, warm = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
with Replicas(Opt(n_replicas=n_rep, n_events=10 ** 9, order="ref", rng="philox")) as s:
    s.run(warm)
    ms = s.run(n_ev)
    print(f"{n_rep} x {n_ev}: {ms:.2f} ms, {n_rep * n_ev / ms / 1e3:.2f} M events/s, {ms * 1e3 / n_ev:.2f} us per event step")
