"""FAST vs REF arithmetic order, FREE-RUNNING (hard part H1).

Channel choice is chaotic in the fp32 summation order: after the first near-tie the FAST order
(the throughput kernel's tree) and the REF order (the Accelerate Interpreter's right fold) pick different
channels and the two simulations drift apart event by event.  Both are legal executions of the same Accelerate
program, so what must agree is the STATISTICS of a run, not its trace.  This tool runs the same seeds through
both device kernels -- dca::k_run<ORDER_REF> and dca::fast::k_run_fast -- and reports, per order, mean +- standard
error over seeds of
    cumulative blocking probability  rejected_new / (arrivals_new + 1)     (statsCumus, src/Stats.hs:83-92)
    final average-reward estimate    avgReward                             (src/Agent.hs:78)
    mean |loss|                      mean over events of |TD error|        (src/Agent.hs:65,80)
and the z-scores of the difference, unpaired (sqrt(SE_f^2 + SE_r^2)) and paired by seed.

  python tools/order_stats.py [n_seeds] [n_events] > profiles/r2_fast_vs_ref_order_stats.md
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIGS = {
    "config 1 (call_rate 200/h, call_dur 3 min, hoff_prob 0)": dict(),
    "config 2 (hoff_prob 0.15, call_dur_hoff 1 min)": dict(hoff_prob=0.15, call_dur_hoff=1.0),
}
METRICS = ("blocking_probability", "avg_reward", "mean_abs_loss")


def run_order(order, n_seeds, n_events, seed, **kw):
    """-> dict metric -> float64[n_seeds], plus the raw summary"""
    from haskelldca_b200 import Opt, Replicas

    with Replicas(Opt(n_replicas=n_seeds, n_events=n_events, order=order, rng="philox", rng_seed=seed, trace=n_events, **kw)) as sims:
        ms = sims.run(n_events)
        s = sims.read_summary()
        assert (s["iter"] == n_events).all(), (order, np.unique(s["status"]))
        st = s["stats"].astype(np.float64)
        mal = np.array([np.abs(sims.read_trace(r)["loss"].astype(np.float64)).mean() for r in range(n_seeds)])
        return {"blocking_probability": st[:, 2] / (st[:, 0] + 1.0), "avg_reward": s["avg_reward"].astype(np.float64),
                "mean_abs_loss": mal, "_summary": s, "_ms": ms}


def compare(n_seeds=256, n_events=10000, seed=0, **kw):
    fast = run_order("fast", n_seeds, n_events, seed, **kw)
    ref = run_order("ref", n_seeds, n_events, seed, **kw)
    rows = {}
    for m in METRICS:
        f, r = fast[m], ref[m]
        se_f, se_r = f.std(ddof=1) / np.sqrt(n_seeds), r.std(ddof=1) / np.sqrt(n_seeds)
        d = f - r
        se_d = d.std(ddof=1) / np.sqrt(n_seeds)
        rows[m] = dict(fast_mean=f.mean(), fast_se=se_f, ref_mean=r.mean(), ref_se=se_r, diff=d.mean(),
                       z_unpaired=(f.mean() - r.mean()) / np.hypot(se_f, se_r), z_paired=d.mean() / se_d if se_d > 0 else 0.0,
                       seeds_identical=int((d == 0).sum()))
    return rows, fast, ref


def oracle_cross_check(ref, n_check, n_events, seed, **kw):
    """the REF-order device runs of the first n_check seeds against the CPU oracle (REF order), bit for bit"""
    import pyoracle as po

    s = ref["_summary"]
    for r in range(n_check):
        o = po.Sim(po.default_opt(order=po.ORDER_REF, rng_mode=po.RNG_PHILOX, seed=seed, replica=r, n_events=n_events, **kw))
        o.run(n_events)
        want = o.read()
        assert np.array_equal(want["stats"], s["stats"][r]), r
        assert np.float32(want["avg_reward"]).view(np.uint32) == s["avg_reward"][r].view(np.uint32), r
    return n_check


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    print(f"# FAST vs REF arithmetic order, free-running: {n_seeds} seeds x {n_events} events per configuration\n")
    print("Same Philox streams (seed 0, replica = seed index) through `dca::fast::k_run_fast` (FAST) and `dca::k_run<ORDER_REF>` (REF) on one B200;")
    print("mean +- standard error over seeds.  z = difference of the means over its standard error (unpaired: independent-sample")
    print("formula; paired: per-seed differences).  |z| < 3 is asserted by `tests/test_gpu_order_stats.py`.\n")
    out = {}
    for name, kw in CONFIGS.items():
        rows, fast, ref = compare(n_seeds, n_events, 0, **kw)
        n_chk = oracle_cross_check(ref, 8, n_events, 0, **kw)
        print(f"## {name}\n")
        print("| metric | FAST mean +- SE | REF mean +- SE | FAST - REF | z unpaired | z paired | seeds with identical value |")
        print("|---|---|---|---|---|---|---|")
        for m, r in rows.items():
            print(f"| {m} | {r['fast_mean']:.6g} +- {r['fast_se']:.3g} | {r['ref_mean']:.6g} +- {r['ref_se']:.3g} | {r['diff']:+.3g} | "
                  f"{r['z_unpaired']:+.2f} | {r['z_paired']:+.2f} | {r['seeds_identical']} / {n_seeds} |")
        print(f"\nREF-order device runs of seeds 0..{n_chk - 1} equal the CPU oracle (REF order) bit for bit (statistics, average reward).")
        print(f"Device time: FAST {fast['_ms']:.0f} ms, REF {ref['_ms']:.0f} ms for {n_seeds} x {n_events} events (trace instantiations).\n")
        out[name] = rows
    print("```json")
    print(json.dumps(out, default=float))
    print("```")


if __name__ == "__main__":
    main()
