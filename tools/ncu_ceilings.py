"""The ceilings that actually bind dca::fast::k_run_fast, from one `ncu --set full` capture (tools/gpu_ncu_full.sh):
issue slots, the shared-memory data pipe, instruction counts and the real DRAM traffic.  Writes the JSON that bench.py
quotes next to the HBM roofline (profiles/kernel_ceilings.json) and prints a one-screen summary.
usage: ncu_ceilings.py report.ncu-rep n_events_in_the_profiled_launch [out.json]"""
import csv, io, json, subprocess, sys

rep, events = sys.argv[1], float(sys.argv[2])
out_path = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
m = {}
for h, u, v in zip(hdr, unit, val):
    try:
        m[h] = float(v.replace(",", ""))
    except ValueError:
        m[h] = v
def g(k, d=None):
    return m.get(k, d)
n_sm = 148
cycles = g("sm__cycles_elapsed.max") or g("sm__cycles_elapsed.avg")
wf = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
stalls = {h.split("issue_stalled_")[1].split("_per_issue_active")[0]: v for h, v in m.items()
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and isinstance(v, float)}
dur_unit = unit[hdr.index("gpu__time_duration.sum")]
dur_ms = g("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(dur_unit, 1.0)
res = {
    "kernel": g("Kernel Name"), "grid": g("Grid Size"), "events_in_launch": events,
    "duration_ms_under_ncu": dur_ms,
    "registers_per_thread": g("launch__registers_per_thread"), "shared_mem_per_cta_kb": g("launch__shared_mem_per_block_dynamic"),
    "ctas_per_sm": g("launch__occupancy_limit_registers"),
    "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "eligible_warps_per_cycle": g("smsp__warps_eligible.avg.per_cycle_active"),
    "warp_instructions_per_event": g("smsp__inst_executed.sum") / events,
    "shared_wavefronts_per_event": wf / events,
    "shared_pipe_pct_of_peak": 100.0 * wf / (cycles * n_sm) if cycles else None,
    "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "pipe_lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "dram_bytes_read": g("dram__bytes_read.sum") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(unit[hdr.index("dram__bytes_read.sum")], 1),
    "dram_bytes_write": g("dram__bytes_write.sum") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(unit[hdr.index("dram__bytes_write.sum")], 1),
    "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "stalls_per_issued_instruction": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8]),
    "source": f"ncu --set full --clock-control none --import-source on -k regex:k_run_fast -s 1 -c 1 python tools/prof_run.py 4096 2000 20000 ({rep.split('/')[-1]})",
}
res["dram_bytes_per_launch"] = res["dram_bytes_read"] + res["dram_bytes_write"]
# the profiled launch runs ONE unit (replica, <= 4096 events) per CTA: a replica's state is staged in and out once per unit
try:
    n_units = int(str(res["grid"]).strip("()").split(",")[0])
    res["units_in_launch"] = n_units
    res["dram_bytes_per_unit"] = res["dram_bytes_per_launch"] / n_units
except Exception:
    pass
print(json.dumps(res, indent=1))
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)
