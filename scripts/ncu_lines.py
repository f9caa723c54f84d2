#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an .ncu-rep (needs -lineinfo builds):
ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python scripts/ncu_lines.py src.csv [top]"""
import csv, os, sys
def num(x):
    try: return int(x)
    except ValueError: return 0
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; out = []; fname = "?"; kernels = 0
for r in rows:
    if r and r[0] == 'File Path': fname = os.path.basename(r[1]); continue
    if r and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or not r or r[0] == '': continue
    try: ln = int(r[0])
    except ValueError: continue
    i_inst = hdr.index('Instructions Executed'); i_s = hdr.index('Warp Stall Sampling (All Samples)')
    out.append((fname, ln, r[1].strip(), num(r[i_inst]), num(r[i_s])))
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total inst {ti}, samples {ts}")
for fn, ln, src, ins, st in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{fn[:16]:16s}{ln:5d} inst {100*ins/ti:5.1f}%  stall {100*st/max(ts,1):5.1f}%  {src[:100]}")
