#!/usr/bin/env python
"""Static instruction footprint of a kernel per call site of its body (outermost source line through the inlining chain
that nvdisasm -gi prints): how many SASS instructions (16 B each) every piece of the kernel costs in the I-cache.

  python tools/sass_footprint.py obj.o kernel_substring [top_n]
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

obj, kname = sys.argv[1:3]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt = Counter()
chain, infn, fresh, total = [], False, True, 0
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infn = kname in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s+/\*([0-9a-f]{4,})\*/", l):
        cnt[chain[-1] if chain else ("?", 0)] += 1
        total += 1
        fresh = True
print(f"{total} instructions = {total * 16 / 1024:.1f} KB")
src = {}
for (f, n), c in cnt.most_common(top_n):
    p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
    if f not in src:
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    line = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    print(f"{f}:{n:<5d} {c:6d} {c * 16 / 1024:6.1f} KB  {line}")
