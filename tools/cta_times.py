"""Development tool: per-CTA run time of the FAST kernel by SM (library built with -DDCA_PHASE_CLOCK).
usage: cta_times.py [n_replicas] [n_events] [warm_events]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haskelldca_b200 import build as B
lib = os.path.join(B.OUT_DIR, "libdca_b200_phase.so")
if not os.path.exists(lib):
    B.build(force=True, extra_flags=["-DDCA_PHASE_CLOCK"], out=lib)
os.environ["DCA_LIB"] = lib
from haskelldca_b200 import Opt, Replicas, _ffi

n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n_ev = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
with Replicas(Opt(n_replicas=n_rep, n_events=warm + n_ev, order="fast", rng="philox")) as s:
    s.run(warm)
    ms = s.run(n_ev)
    out = np.zeros((n_rep, 2), dtype=np.uint64)
    _ffi.load().dca_debug_cta(out.ctypes.data_as(C.c_void_p), n_rep)
    cyc, smid = out[:, 0].astype(np.float64) / n_ev, out[:, 1].astype(int)
    print(f"{n_rep} x {n_ev}: {ms:.2f} ms, {n_rep*n_ev/ms/1e3:.2f} M events/s")
    print(f"cycles per event per CTA: min {cyc.min():.0f}  median {np.median(cyc):.0f}  max {cyc.max():.0f}  (max/median {cyc.max()/np.median(cyc):.3f})")
    per_sm = np.array([cyc[smid == k].mean() for k in range(smid.max() + 1) if (smid == k).any()])
    ids = np.array([k for k in range(smid.max() + 1) if (smid == k).any()])
    order = np.argsort(per_sm)
    print("fastest SMs:", [(int(ids[i]), round(per_sm[i])) for i in order[:8]])
    print("slowest SMs:", [(int(ids[i]), round(per_sm[i])) for i in order[-8:]])
    within = np.mean([cyc[smid == k].std() for k in ids])
    print(f"spread between SMs (std of SM means) {per_sm.std():.0f} cycles, mean spread inside an SM {within:.0f} cycles")
    # by pairs of SMs (TPC) and by blocks of 16 (a GPC-like grouping)
    for g in (2, 16):
        m = np.array([per_sm[(ids // g) == k].mean() for k in range(ids.max() // g + 1) if ((ids // g) == k).any()])
        print(f"means by groups of {g} SM ids: min {m.min():.0f} max {m.max():.0f}")
