"""Print the SASS of the agents' event loop (or the event warp's) of k_run_fast<false> in a built library, one instruction
per line, for reading.  usage: sass_loop.py lib.so [agents|event]"""
import re, subprocess, sys
KERNEL = "_ZN3dca4fast10k_run_fastILb0ELb0EEEvNS_7RunArgsE"
lib = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else "agents"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
body = sass.split("Function : " + KERNEL)[1].split("Function : ")[0]
ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+((?:@!?U?P\d+\s+)?[A-Z][^;]*);", body)
ops = [(a, t.strip()) for a, t in ins]
def idx(pred, start=0):
    for i in range(start, len(ops)):
        if pred(ops[i][1]): return i
b2 = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x2"))
b4 = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x4"))
top = max(i for i in range(b2) if ops[i][1].startswith("BAR.SYNC.DEFER_BLOCKING 0x0"))
b3 = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x3"))
arv4 = idx(lambda t: t.startswith("BAR.ARV 0x4"))
etop = max(i for i in range(b3) if ops[i][1].startswith("BAR.SYNC.DEFER_BLOCKING 0x0"))
eend = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x0"), arv4)
lo, hi = (top, b4 + 4) if which == "agents" else (etop, eend)
for a, t in ops[lo:hi]: print(a, t)
