#!/usr/bin/env python
"""Per-kernel SASS evidence: counts of the Blackwell tensor-core / TMA / TMEM mnemonics in the shipped library.

  python tools/sass_summary.py [libasr_b200.so] > profiles/r02_sass.md

UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA
tensor-map copies), UBLKCP = cp.async.bulk (1-D bulk copies), HMMA = legacy mma.sync, LDSM = ldmatrix,
SYNCS = mbarrier operations, UTCBAR = tcgen05.commit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "asr_transformer_b200", "libasr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "SYNCS"]
counts = collections.OrderedDict()
total = collections.Counter()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        total[cur] += 1
        for k in MNEMONICS:
            if op.startswith(k):
                counts[cur][k] += 1
dem = subprocess.run(["cu++filt"] + list(counts.keys()), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(counts.keys(), dem)) if len(dem) == len(counts) else {k: k for k in counts}


def short(n):
    n = re.sub(r"asr::\(anonymous namespace\)::|asr::<unnamed>::|\(anonymous namespace\)::", "", n)
    i = n.rfind("(")
    return (n[:i] if i > 0 else n).replace("void ", "")


print("# SASS summary of asr_transformer_b200/libasr_b200.so (tools/sass_summary.py; cuobjdump -sass, sm_100a)\n")
print("| kernel | SASS instr | " + " | ".join(MNEMONICS) + " |")
print("|---|---:|" + "---:|" * len(MNEMONICS))
for k, c in counts.items():
    print(f"| `{short(names[k])}` | {total[k]} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
