"""Per-function split of executed warp instructions from an .ncu-rep source page.
usage: ncu_funcs.py report.ncu-rep n_events"""
import csv, subprocess, sys, io, re, collections, os
rep, events = sys.argv[1], float(sys.argv[2])
def num(v):
    try:
        return float(v)
    except ValueError:
        return 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs, cur = [], None
for r in rows:
    if r and r[0] == 'File Path':
        cur = {'file': r[1], 'rows': []}; secs.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
grand = 0
for s in secs:
    hdr = None; lines = []
    for r in s['rows']:
        if r and r[0] == 'Line No':
            hdr = r; iI = hdr.index('Instructions Executed'); iS = hdr.index('Warp Stall Sampling (All Samples)'); continue
        if hdr and r and r[0].isdigit():
            lines.append((int(r[0]), num(r[iI]), num(r[iS])))
    tot = sum(v for _, v, _ in lines); grand += tot
    print(f"{os.path.basename(s['file']):28s} {tot/events:8.1f} instr/event")
    path = s['file'].replace('/tmp/code/tsoernes__haskelldca/repo', os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if not os.path.exists(path) or not path.endswith(('.cuh', '.cu', '.h')):
        continue
    src = open(path).read().split('\n')
    fn_at = {}; curfn = '(top)'
    for i, line in enumerate(src, 1):
        m = re.match(r'^(?:template.*>\s*)?__(?:device|global)__.*?\b(\w+)\s*\(', line)
        if m: curfn = m.group(1)
        fn_at[i] = curfn
    agg = collections.OrderedDict(); smp = collections.Counter()
    for ln, v, sp in lines:
        f = fn_at.get(ln, '?'); agg[f] = agg.get(f, 0) + v; smp[f] += sp
    tots = sum(smp.values()) or 1
    for f, v in sorted(agg.items(), key=lambda x: -x[1]):
        if v / events >= 0.5:
            print(f"      {f:26s} {v/events:8.1f}   stall samples {100*smp[f]/tots:5.1f}%")
print(f"total {grand/events:.0f} instr/event")
