#!/usr/bin/env python
"""Top SASS instructions of an ncu report by stall samples (ncu -i rep --page source --csv), with their neighbours'
opcodes so that a wait site can be recognised.   python tools/ncu_sass_top.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scol = [h for h in hdr if h.startswith("Warp Stall Sampling (All")][0]
icol = [h for h in hdr if h.startswith("Instructions Executed")][0]
data = [r for r in rows[2:] if len(r) > 5]
tot = sum(int(r[ix[scol]] or 0) for r in data)
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix[scol]] or 0))[:top]
print("total samples", tot)
for i in sorted(order):
    r = data[i]
    print(f"{i:5d} {r[ix['Address']][-6:]} samples {int(r[ix[scol]] or 0):5d} ({100*int(r[ix[scol]] or 0)/tot:4.1f}%) exec {r[ix[icol]]:>9}  {r[ix['Source']][:70]}")
