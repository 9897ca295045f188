"""Static SASS opcode summary of the run kernels in the built library (cuobjdump -sass) plus ptxas' register / spill
report: what proves the tensor-memory scratchpad (LDTM / STTM), packed fp32 (FFMA2 / FADD2 / FMUL2), warp reductions
(REDUX / CREDUX), named barriers (BAR.SYNC / BAR.ARV) and the absence of local-memory traffic in the hot loops.
usage: sass_summary.py [lib.so] > profiles/r2_v10_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "haskelldca_b200", "_build", "libdca_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = dict(re.findall(r"Function (\S+):\n\s*(REG:.*)", res))
funcs, cur = collections.OrderedDict(), None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
KEY = ["LDTM", "STTM", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "PRMT", "I2F", "IMAD", "R2UR", "MOV", "LDS", "STS", "SHFL", "REDUX", "CREDUX",
       "VOTE", "BAR", "LDL", "STL", "LDG", "STG", "DADD", "DMUL", "DFMA", "CALL", "BRA"]
print(f"library: {os.path.relpath(lib, ROOT)}\n")
for name, ops in funcs.items():
    if "k_run" not in name and "k_step_fast" not in name:
        continue
    c = collections.Counter(ops)
    base = collections.Counter()
    for op, n in c.items():
        base[op.split(".")[0]] += n
    print(f"{name}\n  {usage.get(name, '')}\n  {len(ops)} instructions")
    print("  " + "  ".join(f"{k} {base[k]}" for k in KEY if base[k]))
    det = {op: n for op, n in c.items() if op.split(".")[0] in ("LDTM", "STTM", "BAR", "REDUX", "CREDUX", "LDL", "STL")}
    print("  detail: " + "  ".join(f"{op} {n}" for op, n in sorted(det.items())) + "\n")
