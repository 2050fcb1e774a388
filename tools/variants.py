#!/usr/bin/env python
"""Build experiment variants of the library: one source compiled with extra -D flags, linked with the stock objects.

  python tools/variants.py decode_cluster.cu name1:-DFOO=1 name2:-DFOO=2,-DBAR ...
-> asr_transformer_b200/build/variants/libasr_<name>.so   (select with ASR_B200_LIB=<path>; ships to the GPU box)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asr_transformer_b200 import build as B  # noqa: E402

B.build()
src = sys.argv[1]
out = os.path.join(B.HERE, "build", "variants")
os.makedirs(out, exist_ok=True)


def one(spec):
    name, _, defs = spec.partition(":")
    flags = [d for d in defs.split(",") if d]
    o = os.path.join(out, f"{src[:-3]}_{name}.o")
    r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas", "-v"] + flags + ["-c", os.path.join(B.CSRC, src), "-o", o],
                       capture_output=True, text=True)
    if r.returncode:
        return name + " FAILED\n" + r.stderr[-3000:]
    objs = [os.path.join(B.HERE, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != src] + [o]
    lib = os.path.join(out, f"libasr_{name}.so")
    subprocess.run([B._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    spills = [l.strip() for l in r.stderr.splitlines() if "spill" in l and " 0 bytes spill stores" not in l]
    return f"{name}: {lib}\n  " + "\n  ".join(spills[:12])


with ThreadPoolExecutor(max_workers=4) as ex:
    for msg in ex.map(one, sys.argv[2:]):
        print(msg)
