#!/usr/bin/env python
"""bench.py -- processed simulator events/sec on B200 (BASELINE.json's metric).

A "step" is one pass of the hot path over one batch of synthetic call traffic:
mkSimState for every replica on the device, then `n_events` events per replica
through the fused kernel (getAction -> environmentStep -> gridStepTrain each).
Workload at any N: BASELINE.json configs[2], 4096 independent replicas x 100 000
events per GPU (weak scaling: replicas are independent, no data-path collective;
the only collective is the NCCL all-reduce of the blocking-probability counters).

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                      CPU arm: the oracle port in the
                                                            reference's own (Interpreter) order
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "processed events/sec (all replicas, device-timed)"
UNIT = "events/s"
ALG_BYTES_PER_EVENT = 4 * 3479 * 4  # read+write of wNet and wGradCorr, SURVEY.md 8(d)


def kernel_ceilings():
    """What one `ncu --set full` capture of k_run_fast says about the limits that actually bind it (issue slots, the
    shared-memory data pipe) and its real DRAM traffic: profiles/kernel_ceilings.json, written by tools/ncu_ceilings.py."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_ceilings.json")) as f:
            return json.load(f)
    except Exception:
        return None


def measured_traffic(n_rep=None, n_ev=None):
    """DRAM bytes of one k_run_fast launch on the bench workload, from the committed ncu capture (dram__bytes_read.sum +
    dram__bytes_write.sum of `ncu --set full`).  A launch stages every replica's state in and out once per UNIT of
    dca_run_unit_events() events; the capture measures launches of one unit per replica, so the bench launch's traffic
    is the captured bytes per unit x the units it runs."""
    c = kernel_ceilings()
    if c and c.get("dram_bytes_per_unit") and n_rep and n_ev:
        from haskelldca_b200 import _ffi

        unit = int(_ffi.load().dca_run_unit_events())
        units = n_rep * -(-n_ev // unit)
        return float(c["dram_bytes_per_unit"]) * units, f"{c['dram_bytes_per_unit']:.0f} B per unit x {units} units of {unit} events; " + c.get("source", "")
    if c and c.get("dram_bytes_per_launch"):
        return float(c["dram_bytes_per_launch"]), c.get("source", "")
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_launch"]), t.get("source", "")
    except Exception:
        return None, "no ncu capture committed"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


_JSON_FD = None


def own_stdout():
    """The contract is ONE JSON line on stdout: libraries that print there on their own (NCCL announces its version on
    stdout when NCCL_DEBUG is set) are sent to stderr -- file descriptor 1 is pointed at 2 and the line goes out through
    a duplicate of the original."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def host_threads():
    """all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not cap the CPU arm)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_leg(order, n_replicas, n_events, threads, optimised=False):
    """Time the CPU restatement (the reference itself needs GHC + LLVM 6 and cannot run here): the literal port
    (oracle/dca_oracle.c: materialised afterstates + dense dots, the reference's formulation) or, with
    `optimised`, the CPU competitor (oracle/dca_cpu_fast.c: delta-form q-values, bit-mask grid, AVX2 -- bit-exact
    against the literal port in FAST order, tests/test_oracle_cpu_fast.py).  OpenMP over replicas either way."""
    import pyoracle as po

    o = po.default_opt(order=order, rng_mode=po.RNG_PHILOX, seed=0, n_events=n_events)
    run = po.run_replicas_fast if optimised else po.run_replicas
    t0 = time.perf_counter()
    total, stats, _, _ = run(o, 0, n_replicas, n_events, threads)
    dt = time.perf_counter() - t0
    return total / dt, dt, total, stats


def oracle_spot_check(sims, first, n_rep, n_ev, hoff_prob):
    """After the timed region: the replicas this rank just ran (k_run_fast<false>, the benchmarked instantiation)
    against their own oracle runs, bit for bit -- one replica of every wave / event-warp rotation of the launch
    (block ids 0, 740, 1480, 2220 on a 148-SM B200) and the last one.  The oracle runs proceed on host threads
    (ctypes releases the GIL).  Returns (ok, replicas checked, first mismatch or None)."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    import pyoracle as po

    spot = sorted({r for r in (0, 740, 1480, 2220, n_rep - 1) if 0 <= r < n_rep})

    def one(r):
        o = po.Sim(po.default_opt(order=po.ORDER_FAST, rng_mode=po.RNG_PHILOX, seed=0, replica=first + r, n_events=n_ev,
                                  hoff_prob=hoff_prob))
        o.run(n_ev)
        return o.read()

    with ThreadPoolExecutor(max_workers=len(spot)) as ex:
        wants = list(ex.map(one, spot))
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    for r, want in zip(spot, wants):
        got = sims.read_state(r)
        for k in ("grid", "frep", "stats"):
            if not np.array_equal(got[k], want[k]):
                return False, [first + x for x in spot], f"replica {first + r}: {k}"
        for k in ("w", "wg", "avg_reward"):
            if not np.array_equal(u32(got[k]), u32(want[k])):
                return False, [first + x for x in spot], f"replica {first + r}: {k}"
        if got["iter"] != want["iter"] or got["event"] != want["event"]:
            return False, [first + x for x in spot], f"replica {first + r}: iter / next event"
    return True, [first + x for x in spot], None


def ref_order_leg(device, first, n_rep=4096, warm=3000, n_ev=1000):
    """Outside the timed region: the same workload through the REFERENCE-ORDER run kernel (dca::k_run<ORDER_REF>: the
    Accelerate Interpreter's arithmetic, every q a dense 3479-term right fold) -- what a `RefOrder` user of dca_run gets.
    Steady state (the 70-candidate transient of the empty grid is over after ~3000 events), device-timed; the first and
    the last replica are compared bit for bit with the CPU oracle in the same order.  Never raises."""
    try:
        import numpy as np
        import pyoracle as po

        from haskelldca_b200 import Opt, Replicas

        with Replicas(Opt(n_replicas=n_rep, n_events=warm + n_ev, order="ref", rng="philox", rng_seed=0, device=device,
                          first_replica=first)) as s:
            s.run(warm)
            ms = s.run(n_ev)
            u32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
            ok = True
            for r in (0, n_rep - 1):
                o = po.Sim(po.default_opt(order=po.ORDER_REF, rng_mode=po.RNG_PHILOX, seed=0, replica=first + r, n_events=warm + n_ev))
                o.run(warm + n_ev)
                want, got = o.read(), s.read_state(r)
                ok = ok and all(np.array_equal(got[k], want[k]) for k in ("grid", "frep", "stats")) and \
                    all(np.array_equal(u32(got[k]), u32(want[k])) for k in ("w", "wg", "avg_reward")) and got["iter"] == want["iter"]
        return {"value": n_rep * n_ev / (ms / 1e3), "unit": UNIT, "per_gpu": True, "kernel": "dca::k_run<ORDER_REF, false>",
                "sample": f"{n_rep} replicas x {n_ev} events after {warm} warm-up events in the same order, device-timed ({ms:.1f} ms)",
                "oracle_check": bool(ok), "oracle_check_what": f"replicas {first} and {first + n_rep - 1} after {warm + n_ev} events: grid, frep, "
                "stats, wNet, wGradCorr, avgReward bit-exact vs the CPU oracle in REF order"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def inlib_allreduce_check(n_dev):
    """dca_allreduce_stats (the in-library NCCL path: one process, one handle per device, ncclCommInitAll) on the first
    min(n_dev, 4) devices of the box: its int64[10] must equal the sum of the per-device dca_sum_stats and the counters
    of the same replicas run on ONE device.  Never raises: the outcome is reported in checks."""
    try:
        import numpy as np

        from haskelldca_b200 import Opt, Replicas, _ffi, allreduce_stats

        n_dev = min(n_dev, _ffi.load().dca_device_count(), 4)
        if n_dev < 2:
            return {"ran": False, "why": "needs two CUDA devices"}
        n_rep, n_ev = 64, 500
        sets, totals = [], []
        try:
            for d in range(n_dev):
                s = Replicas(Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=4, device=d,
                                 first_replica=d * n_rep, hoff_prob=0.1))
                s.run(n_ev)
                sets.append(s)
                totals.append(s.sum_stats()[0])
            out = allreduce_stats(sets)
            again = allreduce_stats(sets)  # the cached communicators serve a second call
        finally:
            for s in sets:
                s.close()
        with Replicas(Opt(n_replicas=n_dev * n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=4, hoff_prob=0.1)) as one:
            one.run(n_ev)
            single = one.sum_stats()[0]
        ok = bool(np.array_equal(out, sum(totals)) and np.array_equal(out, single) and np.array_equal(out, again)
                  and out[8] == n_dev * n_rep * n_ev and out[9] == 0)
        return {"ran": True, "ok": ok, "devices": n_dev, "events": int(out[8]),
                "what": "dca_allreduce_stats == sum of per-device dca_sum_stats == the same replicas on one device"}
    except Exception as e:  # noqa: BLE001
        return {"ran": True, "ok": False, "error": repr(e)[:300]}


def run_reference(args, rank, world):
    """--impl reference: the reference's algorithm on the host cores.  haskelldca is Haskell/Accelerate
    (no GHC, no LLVM 6 here), so this arm runs the oracle port in the reference's own arithmetic
    (ORDER_REF: materialised afterstates + dense n x 3479 GEMV, Interpreter fold order), one simulator
    per host thread."""
    if rank != 0:
        return
    import pyoracle as po

    threads = host_threads()
    n_rep, n_ev = 2 * threads, args.ref_events
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, total, _ = cpu_leg(po.ORDER_REF, n_rep, n_ev, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
    sample = f"{n_rep} replicas x {n_ev} events per step (bounded sample of the 4096 x 100000 workload), {threads} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[2]: 4096 replicas x 100k events, 7x7 grid, 70 channels, call_rate 200/h, call_dur 3 min, hoff_prob 0",
                   "sample": sample, "order": "reference (Accelerate Interpreter fold order, dense GEMV)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=4096, help="replicas per GPU")
    ap.add_argument("--events", type=int, default=100000, help="events per replica per step")
    ap.add_argument("--hoff_prob", type=float, default=0.0)
    ap.add_argument("--ref-events", type=int, default=100000, dest="ref_events")
    ap.add_argument("--cpu-events", type=int, default=100000, dest="cpu_events")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-spot-check", action="store_true", help="skip the oracle spot check of the timed configuration")
    ap.add_argument("--no-ref-order", action="store_true", dest="no_ref_order",
                    help="skip the (untimed-region) throughput figure of the reference-order run kernel")
    ap.add_argument("--no-inlib-allreduce", action="store_true", dest="no_inlib_allreduce",
                    help="N > 1: skip the check of dca_allreduce_stats (in-library NCCL, one process over several devices)")
    args = ap.parse_args()
    own_stdout()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from haskelldca_b200 import Opt, Replicas, _ffi
    from haskelldca_b200.replica_shards import shard

    _ffi.load(build_if_missing=False)  # the prebuilt sm_100a library, or fail
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    # the in-library NCCL reduction (single process, several devices) is exercised by rank 0 before it joins the
    # process group -- the other ranks are still in the rendezvous and hold no kernels on their devices
    inlib = inlib_allreduce_check(world) if (world > 1 and rank == 0 and not args.no_inlib_allreduce) else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_ev = args.events
    first, n_rep = shard(args.replicas * world, world, rank)  # weak scaling: args.replicas per GPU, contiguous blocks
    opt = Opt(n_replicas=n_rep, n_events=n_ev, order="fast", rng="philox", rng_seed=0, device=local_rank,
              first_replica=first, hoff_prob=args.hoff_prob)
    sims = Replicas(opt)
    state_mb = n_rep * 50960 / 1e6

    # per-replica parameter arrays = the step's host inputs for the e2e leg (BASELINE config defaults)
    host_params = dict(call_rate=np.full(n_rep, 200.0), call_dur=np.full(n_rep, 3.0), call_dur_hoff=np.full(n_rep, 1.0),
                       hoff_prob=np.full(n_rep, args.hoff_prob), learning_rate=np.full(n_rep, 2.52e-6, np.float32),
                       learning_rate_avg=np.full(n_rep, 0.06, np.float32), learning_rate_grad=np.full(n_rep, 5e-6, np.float32))

    def step():
        """device-resident step: inputs (parameters, seeds) already in HBM"""
        sims.reset()
        ms = sims.elapsed_ms()
        ms += sims.run(n_ev)
        return ms

    agents_out = None

    def step_e2e(with_agents=False):
        """the same step through the C ABI with HOST buffers: H2D of the parameters, run, D2H of the per-replica results
        (status, iter, average reward, the 8 Stats counters) -- and, with_agents, of every replica's learned wNet and
        wGradCorr (2 x 3479 floats per replica)"""
        nonlocal agents_out
        h2d = sims.set_params(**host_params)
        sims.reset()
        sims.run(n_ev)
        summ = sims.read_summary()
        d2h = sum(v.nbytes for v in summ.values())
        if with_agents:
            agents_out = sims.read_agents(agents_out)
            d2h += agents_out[0].nbytes + agents_out[1].nbytes + agents_out[2].nbytes
        return h2d, d2h, summ

    for _ in range(args.warmup):
        step()
    launches0 = sims.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    run_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step()
        run_ms += sims.elapsed_ms()  # the run kernel alone (last launch of the step)
    barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = sims.launch_count() - launches0

    total, dev_ptr = sims.sum_stats()
    events_per_step = int(total[8])  # events actually processed by this rank in the last step
    ok_local = int(total[9]) == 0 and events_per_step == n_rep * n_ev

    # max over ranks of the device time; sum over ranks of the events; NCCL all-reduce of the stats counters
    t = torch.tensor([dev_ms, run_ms, wall_s * 1e3], dtype=torch.float64, device="cuda")
    st = torch.from_numpy(total.copy()).cuda()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(st, op=dist.ReduceOp.SUM)  # blocking-probability statistics over NVLink
    dev_ms_max, run_ms_max, wall_ms_max = t.tolist()
    st = st.cpu().numpy()
    events_all = int(st[8])
    value = events_all * args.steps / (dev_ms_max / 1e3)

    # end-to-end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            h2d, d2h, summ = step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t1
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": events_all * args.steps / te.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "reads_back": "per-replica results: status, iter, average reward, the 8 Stats counters (blocking probabilities)"}
        # variant: the learned weights of every replica cross PCIe too
        step_e2e(True)
        barrier()
        t2 = time.perf_counter()
        for _ in range(args.steps):
            h2d_w, d2h_w, _ = step_e2e(True)
        barrier()
        tw = torch.tensor([time.perf_counter() - t2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e["with_learned_weights"] = {"value": events_all * args.steps / tw.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d_w),
                                       "d2h_bytes_per_step": int(d2h_w),
                                       "reads_back": "the above plus wNet and wGradCorr of every replica (dca_read_agents)"}

    # the timed configuration against the oracle (every rank checks replicas of its own shard; the state on the
    # device is the one the last step left behind)
    spot = None
    if not args.no_spot_check:
        ok_spot, spot_reps, spot_msg = oracle_spot_check(sims, first, n_rep, n_ev, args.hoff_prob)
        tk = torch.tensor([1 if ok_spot else 0, len(spot_reps)], dtype=torch.int64, device="cuda")
        if world > 1:
            tmin = tk[:1].clone()
            dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            tsum = tk[1:].clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            tk = torch.cat([tmin, tsum])
        spot = {"ok": bool(tk[0].item()), "replicas_checked_all_ranks": int(tk[1].item()), "rank0_replicas": spot_reps,
                "rank0_mismatch": spot_msg, "events": n_ev,
                "what": "grid, frep, stats, wNet, wGradCorr, avgReward, iter, next event bit-exact vs the CPU oracle (FAST order)"}

    ref_order = ref_order_leg(local_rank, first) if (rank == 0 and not args.no_ref_order) else None
    barrier()  # (the other ranks wait for rank 0's reference-order leg before the process group is torn down)
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        traffic, traffic_src = measured_traffic(n_rep, n_ev)
        ceil = kernel_ceilings()
        binding = None
        if ceil:
            binding = {k: ceil.get(k) for k in ("issue_active_pct", "shared_pipe_pct_of_peak", "warp_instructions_per_event",
                                                "shared_wavefronts_per_event", "eligible_warps_per_cycle", "pipe_alu_pct", "pipe_fma_pct",
                                                "pipe_lsu_pct", "dram_throughput_pct", "registers_per_thread", "shared_mem_per_cta_kb",
                                                "ctas_per_sm", "stalls_per_issued_instruction", "source")}
            if ceil.get("warp_instructions_per_event"):
                # 4 schedulers x 148 SMs x the SM clock issue slots per second: what 100 % issue utilisation would give
                binding["events_per_s_at_full_issue_rate"] = 4 * 148 * 1.965e9 / ceil["warp_instructions_per_event"]
        per_gpu_run = n_rep * n_ev * args.steps / (run_ms_max / 1e3)  # the dominant kernel, per GPU
        achieved = per_gpu_run * ALG_BYTES_PER_EVENT / 1e9
        bp = float(st[2]) / (float(st[0]) + 1.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: {n_rep} replicas/GPU x {n_ev} events, 7x7 grid, 70 channels, call_rate 200/h, "
                                   f"call_dur 3 min, hoff_prob {args.hoff_prob}, lr 2.52e-6/0.06/5e-6, Philox seed 0",
                       "replicas_per_gpu": n_rep, "events_per_replica": n_ev, "order": "fast", "rng": "philox",
                       "l2": f"inputs larger than L2 ({state_mb:.0f} MB of replica state per GPU vs 126 MB L2; state is re-initialised every step)",
                       "parallelism": f"{world} x independent replica shards, NCCL all-reduce of stats only"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "dca::fast::k_run_fast<false>", "peak_source": peak_src, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": n_rep * n_ev * ALG_BYTES_PER_EVENT,
                         "regime": "on-chip residency: NOT an HBM utilisation figure",
                         "binding_ceilings": binding,
                         "note": "achieved = events/s/GPU x 55 664 algorithmic bytes (r+w of wNet, wGradCorr per event, Agent.hs:67-76) over the "
                                 "kernel's own CUDA-event duration, against the measured HBM peak as SURVEY.md 8(d) defines the fraction.  The "
                                 "kernel keeps a replica's weights on chip (shared memory + tensor memory) for the whole launch, so real DRAM "
                                 "traffic (`traffic`) is one state blob in and out per replica per unit of 4096 events and frac > 1 is multi-event residency, "
                                 "not bandwidth; what binds is in `binding_ceilings` (ncu: issue slots, the shared-memory data pipe, stalls)"},
            "e2e": e2e,
            "reference_order": ref_order,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "checks": {"all_replicas_success": bool(ok_local), "blocking_probability": bp, "events_per_step": events_all,
                       "wall_ms_per_step": wall_ms_max / args.steps,
                       "oracle_spot_check": spot["ok"] if spot else None, "oracle_spot_check_detail": spot,
                       "nccl_stats_allreduce": {"torch_distributed_ranks": world, "summed_events": events_all,
                                                "in_library_dca_allreduce_stats": inlib}},
        }
        if not args.no_cpu_baseline and world == 1:
            import pyoracle as po

            threads = host_threads()
            # the fair CPU competitor: delta-form, bit-mask, AVX2, OpenMP over replicas (bit-exact vs the literal port)
            n_cpu_rep = 32 * threads
            v, dt, tot, _ = cpu_leg(po.ORDER_FAST, n_cpu_rep, args.cpu_events, threads, optimised=True)
            # ... and the literal dense port of the reference's formulation (what `--impl reference` times), shorter
            n_lit_rep, n_lit_ev = 2 * threads, args.cpu_events
            v_lit, dt_lit, _, _ = cpu_leg(po.ORDER_REF, n_lit_rep, n_lit_ev, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port-optimised",
                "sample": f"{n_cpu_rep} replicas x {args.cpu_events} events of the same workload (oracle/dca_cpu_fast.c: FAST order, "
                          f"delta-form q-values, bit-mask grid, AVX2; {threads} OpenMP threads, {dt:.1f} s)",
                "gpu_over_cpu": value / v,
                "literal_port": {"value": v_lit, "kind": "port", "sample": f"{n_lit_rep} replicas x {n_lit_ev} events (oracle/dca_oracle.c, REF order: "
                                 f"materialised afterstates + dense n x 3479 dots as the reference formulates them; {threads} threads, {dt_lit:.1f} s)",
                                 "gpu_over_cpu": value / v_lit},
                "note": "both ratios scale with the host's core count (the GPU box of one run had 16 threads, of another 32): quote them with `cores`"}
        emit(line)
    sims.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
