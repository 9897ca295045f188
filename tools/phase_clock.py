"""Development tool: per-role, per-phase cycle counts of the FAST kernel (library built with -DDCA_PHASE_CLOCK).
usage: phase_clock.py [n_replicas] [n_events] [warm_events]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haskelldca_b200 import build as B
lib = os.path.join(B.OUT_DIR, "libdca_b200_phase.so")
if not os.path.exists(lib):
    B.build(force=True, extra_flags=["-DDCA_PHASE_CLOCK"], out=lib)
os.environ["DCA_LIB"] = lib
from haskelldca_b200 import Opt, Replicas, _ffi

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
with Replicas(Opt(n_replicas=n_rep, n_events=warm + n_ev, order="fast", rng="philox")) as s:
    s.run(warm)
    out = (C.c_ulonglong * 48)()
    h = _ffi.load()
    h.dca_debug_phase(out, 1)
    ms = s.run(n_ev)
    h.dca_debug_phase(out, 1)
    print(f"{n_rep} x {n_ev}: {ms:.2f} ms, {n_rep*n_ev/ms/1e3:.2f} M events/s")
    names = ["eval", "wait B2", "decide", "update", "wait B1", "total"]
    for role in range(4):
        v = [out[role * 8 + i] / (n_rep * n_ev) for i in range(6)]
        print(("agent %d" % role if role < 3 else "event  "), "  ".join(f"{n} {x:7.0f}" for n, x in zip(names, v)))
    subs = ["post", "rescans", "masks", "wait cnt2", "candidates", "stop rules", "push next"]
    print("event warp, update phase:", "  ".join(f"{n} {out[4 * 8 + i] / (n_rep * n_ev):6.0f}" for i, n in enumerate(subs)))
