"""Static schedule length of a SASS region: decodes the control word of every instruction (stall count, yield,
scoreboard set / wait) from `cuobjdump -sass` output and sums the compiler-encoded issue stalls.
usage: sass_sched.py file.sass first_line last_line [-v]"""
import re, sys
path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
verbose = len(sys.argv) > 4
lines = open(path).read().split("\n")[lo - 1:hi]
ins = []
i = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* 0x([0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* 0x([0-9a-f]+) \*/", lines[i + 1])
        if m2:
            ins.append((m.group(1), m.group(2).strip(), int(m2.group(1), 16)))
            i += 2
            continue
    i += 1
tot = 0
byop = {}
for addr, txt, hiw in ins:
    stall = (hiw >> 41) & 0xF
    yld = (hiw >> 45) & 1
    wr = (hiw >> 46) & 7
    rd = (hiw >> 49) & 7
    wait = (hiw >> 52) & 0x3F
    tot += max(stall, 1)
    op = txt.split()[1] if txt.startswith("@") else txt.split()[0]
    op = op.split(".")[0]
    c = byop.setdefault(op, [0, 0])
    c[0] += 1; c[1] += max(stall, 1)
    if verbose:
        print(f"{addr} st={stall:2d} y={yld} wr={wr if wr != 7 else '-'} rd={rd if rd != 7 else '-'} wait={wait:06b}  {txt}")
print(f"{len(ins)} instructions, encoded issue cycles {tot}")
for op, (n, c) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"  {op:10s} n={n:4d} cycles={c:5d}  ({c/n:.1f}/instr)")
