"""Development tool: cycles per phase of the REF-order run kernel (library built with -DDCA_PHASE_CLOCK).
usage: ref_phase_clock.py [n_replicas] [n_events] [warm_events]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haskelldca_b200 import build as B
lib = os.path.join(B.OUT_DIR, "libdca_b200_phase.so")
if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in B.deps()):
    B.build(force=True, extra_flags=["-DDCA_PHASE_CLOCK"], out=lib)
os.environ["DCA_LIB"] = lib
from haskelldca_b200 import Opt, Replicas, _ffi

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
with Replicas(Opt(n_replicas=n_rep, n_events=warm + n_ev, order="ref", rng="philox")) as s:
    s.run(warm)
    out = (C.c_ulonglong * 8)()
    h = _ffi.load()
    h.dca_debug_ref_phase(out, 1)
    ms = s.run(n_ev)
    h.dca_debug_ref_phase(out, 1)
    print(f"ref {n_rep} x {n_ev}: {ms:.2f} ms, {n_rep*n_ev/ms/1e3:.2f} M events/s")
    # (thread 0 stamps: it sits on warp 0, which folds, decides, then runs env_post / stop rules / next candidates and its
    #  share of the update; env_ahead runs on warp 3 during the folds)
    names = ["folds (|| env_ahead)", "decision", "env_post + stop rules + next candidates + update"]
    v = [out[i] / (n_rep * n_ev) for i in (1, 2, 3)]
    print("cycles per event:", "  ".join(f"{n} {x:7.0f}" for n, x in zip(names, v)), f"  sum {sum(v):7.0f}")
