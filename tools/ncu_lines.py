"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu_lines.py export.csv [top] [events]   (events: divide instruction counts by this to get warp-instr/event)"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
events = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows[:10]) if 'Source' in r)
hdr = rows[hi]
iS = hdr.index('Warp Stall Sampling (All Samples)'); iI = hdr.index('Instructions Executed')
lines = []
for r in rows[hi + 1:]:
    if r[0] != '' and r[0] != 'Line No':
        try:
            lines.append((int(r[0]), r[1], float(r[iS] or 0), float(r[iI] or 0)))
        except ValueError:
            pass
totS = sum(l[2] for l in lines); totI = sum(l[3] for l in lines)
print(f"total samples {totS:.0f}, total warp-instructions {totI:.3e}" + (f" = {totI/events:.0f} per event" if events else ""))
print("--- by stall samples")
for ln, src, s, i in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{ln:5d} {100*s/totS:5.1f}% smp {100*i/totI:5.1f}% inst  {src[:110]}")
print("--- by instructions")
for ln, src, s, i in sorted(lines, key=lambda x: -x[3])[:top]:
    extra = f" {i/events:7.1f}/ev" if events else ""
    print(f"{ln:5d} {100*i/totI:5.1f}% inst{extra} {100*s/totS:5.1f}% smp  {src[:110]}")
