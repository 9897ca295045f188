"""Build development variants of libdca_b200.so next to the product library (haskelldca_b200/_build/libdca_b200_var_<name>.so)
and print what ptxas made of the throughput kernel: registers, spills, static instruction count, local-memory accesses.
tools/gpu_kernel_exp.sh then times every variant on the GPU box.
usage: build_variants.py name=-DFLAG[,-DFLAG2] ...      (name 'base' with no flags = the product configuration)"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from haskelldca_b200 import build as B

KERNEL = "_ZN3dca4fast10k_run_fastILb0ELb0EEEvNS_7RunArgsE"
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    flags = [f for f in flags.split(",") if f]
    out = os.path.join(B.OUT_DIR, f"libdca_b200_var_{name}.so")
    B.build(force=True, extra_flags=flags, out=out)
    res = subprocess.run(["cuobjdump", "-res-usage", out], capture_output=True, text=True).stdout
    m = re.search(r"Function %s:\n\s*(REG:\S+ STACK:\S+)" % KERNEL, res)
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    body = sass.split("Function : " + KERNEL)[1].split("Function : ")[0]
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", body)
    n = lambda k: sum(1 for o in ops if o == k)
    print(f"{name:24s} {' '.join(flags):40s} {m.group(1) if m else '?'}  instr {len(ops)}  LDL {n('LDL')} STL {n('STL')} "
          f"IMAD {n('IMAD')} LDTM {n('LDTM')} STTM {n('STTM')}", flush=True)
