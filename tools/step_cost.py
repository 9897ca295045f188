"""Cost of the STEP-WISE path (what `runStep`, src/SimRunner.hs:60-97, pays per event when only the two array
functions accFn1 / accFn2 run on the GPU), and the batched throughput of the reference-order run kernel.

  python tools/step_cost.py [calls] [ref_replicas] [ref_events]

Prints one JSON object:
  us_per_call: dca_acc_get_action / dca_acc_grid_step_train (pure operators on HOST arrays, both orders),
               dca_get_action / dca_grid_step_train (device-resident state, both orders) -- wall clock per call
               through the C ABI, median of `calls` calls on a warmed-up state;
  ref_run:     dca::k_run<ORDER_REF> on ref_replicas x ref_events (device time, events/s), next to the FAST kernel
               on the same batch.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from haskelldca_b200 import Opt, Replicas, agent  # noqa: E402


def med_us(fn, n):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return round(1e6 * ts[len(ts) // 2], 1)


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    ref_rep = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    ref_ev = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    out = {"us_per_call": {}, "calls": calls}
    # a warmed-up state: 3000 events of the fused kernel, then its arrays on the host
    with Replicas(Opt(n_events=10 ** 9, order="fast", rng="philox", rng_seed=1)) as sims:
        sims.run(3000)
        st = sims.read_state(0)
    g, f, w, wg, avg = st["grid"], st["frep"], st["w"], st["wg"], float(st["avg_reward"])
    cell = (3, 3)
    for order in ("ref", "fast"):
        ch, nf = agent.get_action(cell, False, w, g, f, order)  # first call allocates the per-device scratch
        out["us_per_call"][f"dca_acc_get_action[{order}]"] = med_us(lambda: agent.get_action(cell, False, w, g, f, order), calls)
        out["us_per_call"][f"dca_acc_grid_step_train[{order}]"] = med_us(
            lambda: agent.grid_step_train(2.52e-6, 0.06, 5e-6, False, cell, ch, f, nf, g, (w, wg, avg), order), calls)
        with Replicas(Opt(n_events=10 ** 9, order=order, rng="philox", rng_seed=1)) as sims:
            sims.run(1000 if order == "ref" else 3000)
            sims.get_action(0, cell, False)
            out["us_per_call"][f"dca_get_action[{order}]"] = med_us(lambda: sims.get_action(0, cell, False), calls)

            def pair():  # a NEW followed by the END of the same call: the state stays where it is
                c = sims.get_action(0, cell, False)
                sims.grid_step_train(0, False, cell, c)
                sims.grid_step_train(0, True, cell, c)

            t = med_us(pair, calls // 3)
            out["us_per_call"][f"dca_grid_step_train[{order}]"] = round((t - out["us_per_call"][f"dca_get_action[{order}]"]) / 2, 1)
    # the reference-order run kernel on a batch (what a `RefOrder` user of dca_run gets)
    res = {}
    for order in ("ref", "fast"):
        with Replicas(Opt(n_replicas=ref_rep, n_events=ref_ev, order=order, rng="philox", rng_seed=0)) as sims:
            ms = sims.run(ref_ev)
            tot = sims.sum_stats()[0]
            assert tot[8] == ref_rep * ref_ev and tot[9] == 0
            res[order] = {"ms": round(ms, 2), "events_per_s": round(ref_rep * ref_ev / ms * 1e3)}
    out["ref_run"] = {"replicas": ref_rep, "events": ref_ev, **res}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
