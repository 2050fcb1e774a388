#!/usr/bin/env python
"""Attribute an ncu SASS source page to the OUTERMOST source line of the kernel body (through the inlining chain that
nvdisasm -gi prints): per call site in the kernel, executed warp instructions, stall samples and the top stall reasons.

  python tools/ncu_regions.py rep.ncu-rep obj.o kernel_substring [top_n] [--inner]

kernel_substring matches the MANGLED name in the object (e.g. "ShapeILi4ELi256ELi64ELi8E" picks one instantiation of
dec_cluster_kernel); a substring that matches another instantiation silently mis-attributes the samples.

--inner additionally splits every call site by the innermost line (helper level).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

args = [a for a in sys.argv[1:] if not a.startswith("--")]
inner = "--inner" in sys.argv
rep, obj, kname = args[:3]
top_n = int(args[3]) if len(args) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 5]
base = min(int(r[ix["Address"]], 16) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
where = {}
chain = []
infn = False
fresh = True
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infn = kname in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        where[int(m.group(1), 16)] = (chain[-1], chain[0]) if chain else (("?", 0), ("?", 0))
        fresh = True
agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
ti = ts = 0.0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    outer, inn = where.get(off, (("?", 0), ("?", 0)))
    key = (outer, inn) if inner else (outer,)
    try:
        i, s = float(r[ix["Instructions Executed"]]), float(r[ix["# Samples"]])
    except ValueError:
        continue
    a = agg[key]
    a[0] += i
    a[1] += s
    for c in stall_cols:
        try:
            a[2][c] += float(r[ix[c]])
        except ValueError:
            pass
    ti += i
    ts += s
src_cache = {}


def src(key):
    f, n = key
    p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
    if not os.path.exists(p):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "asr_transformer_b200", "csrc", f)
    if f not in src_cache:
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[n - 1].strip()[:70] if 0 < n <= len(L) else ""


print(f"total warp instructions {ti:.3e}, samples {ts:.0f}")
print("outer line | instr% | samples% | top stalls | source")
order = sorted(agg.items(), key=(lambda kv: kv[0]) if "--by-line" in sys.argv else (lambda kv: -kv[1][1]))
for key, (i, s, st) in order[:top_n]:
    tops = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    tstr = " ".join(f"{k[6:]}:{100 * v / max(s, 1):.0f}" for k, v in tops)
    name = f"{key[0][0]}:{key[0][1]}" + (f" <- {key[1][0]}:{key[1][1]}" if inner else "")
    print(f"{name:<28s} {100 * i / ti:6.2f} {100 * s / ts:6.2f}  {tstr:<40s} {src(key[-1] if inner else key[0])}")
