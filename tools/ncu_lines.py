#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i rep --page source --csv) with nvdisasm --print-line-info of the cubin:
per CUDA source line, executed warp instructions and stall samples.

  python tools/ncu_lines.py rep.ncu-rep obj.o kernel_substring [top_n]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, kname = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 5]
base = min(int(r[ix["Address"]], 16) for r in data)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# walk the function's text section
line_of = {}
cur = None
infn = False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infn = kname in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0.0, 0.0])
ti = ts = 0.0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    try:
        i, s = float(r[ix["Instructions Executed"]]), float(r[ix["# Samples"]])
    except ValueError:
        continue
    agg[key][0] += i
    agg[key][1] += s
    ti += i
    ts += s
src_cache = {}
def src(key):
    f, n = key
    p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
    if f not in src_cache:
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
print(f"total warp instructions {ti:.3e}, samples {ts:.0f}")
print("line  instr%  samples%  source")
for key, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print(f"{key[0]}:{key[1]:<5d} {100*i/ti:6.2f} {100*s/ts:6.2f}  {src(key)}")
