"""Dump the per-event trace of a --fixed_rng run (REF order, MT19937-64 seed 0) as CSV, one line per event:
iter, time, type, row, col, end_ch, n_candidates, chosen_ch, reward, loss, avg_reward, grid_hash, frep_hash.

Purpose: a maintainer with GHC 8.4 can print the same columns from the reference's
runStep (src/SimRunner.hs:60-97) and diff the two files; the first differing line localises any disagreement
between the reference binary and this implementation / its oracle.  `--oracle` prints the CPU oracle's trace
instead of the GPU's (they are bit-identical; tests/test_gpu_parity.py)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

TYPES = {0: "NEW", 1: "END", 2: "HOFF"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_events", type=int, default=10000)
    ap.add_argument("--hoff_prob", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--order", choices=["ref", "fast"], default="ref")
    ap.add_argument("--oracle", action="store_true", help="CPU oracle instead of the GPU library")
    a = ap.parse_args()
    if a.oracle:
        import pyoracle as po

        o = po.Sim(po.default_opt(order=po.ORDER_REF if a.order == "ref" else po.ORDER_FAST, rng_mode=po.RNG_MT, seed=a.seed,
                                  n_events=a.n_events, hoff_prob=a.hoff_prob))
        _, _, tr = o.run(a.n_events, trace=True)
    else:
        from haskelldca_b200 import Opt, Replicas

        with Replicas(Opt(n_events=a.n_events, order=a.order, rng="mt19937", rng_seed=a.seed, hoff_prob=a.hoff_prob, trace=a.n_events)) as s:
            s.run(a.n_events)
            tr = s.read_trace(0)
    print("iter,time,type,row,col,end_ch,n_candidates,chosen_ch,reward,loss,avg_reward,grid_hash,frep_hash")
    for i, e in enumerate(tr):
        print(f"{i},{float(e['time'])!r},{TYPES[int(e['etype'])]},{int(e['cell']) // 7},{int(e['cell']) % 7},{int(e['end_ch'])},"
              f"{int(e['ncand'])},{int(e['ch'])},{int(e['reward'])},{float(e['loss'])!r},{float(e['avg_reward'])!r},"
              f"{int(e['grid_hash']):016x},{int(e['frep_hash']):016x}")


if __name__ == "__main__":
    main()
