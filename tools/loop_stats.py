"""Static size of the two hot loops of k_run_fast<false> in a built library: instructions between the CTA barrier in front of
the agents' event loop and their barrier 4, and between the event warp's loop head and its back edge (barrier 3 .. ARV 4 region),
with the count of register moves / S2R / NOP -- what ptxas' register allocation added on top of the arithmetic.
usage: loop_stats.py lib.so [...]"""
import collections, re, subprocess, sys
KERNEL = "_ZN3dca4fast10k_run_fastILb0ELb0EEEvNS_7RunArgsE"
for lib in sys.argv[1:]:
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    body = sass.split("Function : " + KERNEL)[1].split("Function : ")[0]
    ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+((?:@!?U?P\d+\s+)?[A-Z][^;]*);", body)
    ops = [(int(a, 16), t.strip()) for a, t in ins]
    def idx(pred, start=0):
        for i in range(start, len(ops)):
            if pred(ops[i][1]): return i
        return None
    b2 = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x2"))
    b4 = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x4"))
    top = max(i for i in range(b2) if ops[i][1].startswith("BAR.SYNC.DEFER_BLOCKING 0x0"))
    b3 = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x3"))
    arv4 = idx(lambda t: t.startswith("BAR.ARV 0x4"))
    etop = max(i for i in range(b3) if ops[i][1].startswith("BAR.SYNC.DEFER_BLOCKING 0x0"))
    eend = idx(lambda t: t.startswith("BAR.SYNC.DEFER_BLOCKING 0x0"), arv4)
    def stats(lo, hi):
        c = collections.Counter()
        for _, t in ops[lo:hi]:
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            if op.startswith("IMAD.MOV") or op == "MOV": op = "MOV*"
            c[op.split(".")[0] if op != "MOV*" else op] += 1
        return hi - lo, c
    na, ca = stats(top, b4)
    ne, ce = stats(etop, eend)
    f = lambda c: " ".join(f"{k} {c[k]}" for k in ("MOV*", "S2R", "NOP", "R2UR", "LDL", "STL", "LDS", "STS", "PRMT", "IMAD"))
    print(f"{lib.split('/')[-1]:36s} agents {na:4d} [{f(ca)}]  event {ne:4d} [{f(ce)}]")
