"""Key metrics of one kernel from an .ncu-rep (raw page) + per-function instruction split (source page).
usage: ncu_summary.py report.ncu-rep [n_events]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; events = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
m = {h: (val[i], unit[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k][0]:>16s} {m[k][1]}")
if events and "smsp__inst_executed.sum" in m:
    print(f"warp instructions / event: {float(m['smsp__inst_executed.sum'][0].replace(',', ''))/events:.0f}")
st = [(h, float(v[0].replace(',', ''))) for h, v in m.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
if not st:
    st = [(h, float(v[0].replace(',', '') or 0)) for h, v in m.items() if h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio")]
tot = sum(v for _, v in st) or 1
print("--- stall reasons (warp-latency per issued instruction)")
for h, v in sorted(st, key=lambda x: -x[1])[:10]:
    print(f"   {h.split('issue_stalled_')[1].split('.')[0]:32s} {v:8.3f}  {100*v/tot:5.1f}%")
