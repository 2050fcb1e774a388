#!/usr/bin/env python
"""Decode (128 utterances) with the NEXT group's front-end + encoder + cross-K/V on a second stream, as the serving loop
runs them; toggles: rotating inputs, stream priority, fresh workspaces."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asr_transformer_b200 import workloads as O  # noqa: E402  (workload registry + synthetic inputs)
from asr_transformer_b200.workloads import build_model  # noqa: E402

cfg = O.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = build_model(cfg, dev)
B = 128
specs = [O.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=1 + i).to(dev) for i in range(4)]
eng = m._eng()


def run(rotate, prio, alt_ws, n=6):
    enc_s = torch.cuda.Stream(dev)
    dec_s = torch.cuda.Stream(dev, priority=-1 if prio else 0)
    ctx = eng.decode_greedy(eng.encode(specs[0], ws_tag="e"), ws_tag="d0", phase="prepare")
    torch.cuda.synchronize()
    durs = []
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(dec_s):
            e0.record()
            eng.decode_greedy(None, phase=ctx)
            e1.record()
        with torch.cuda.stream(enc_s):
            x = specs[(i + 1) % 4] if rotate else specs[0]
            enc = eng.encode(x, ws_tag="e")
            nxt = eng.decode_greedy(enc, ws_tag=("d%d" % ((i + 1) & 1)) if alt_ws else "d1", phase="prepare")
        torch.cuda.synchronize()
        durs.append(e0.elapsed_time(e1))
        ctx = nxt
    return durs


for rotate in (0, 1):
    for prio in (0, 1):
        d = run(rotate, prio, 1)
        print(f"rotate={rotate} priority={prio}: decode ms {' '.join('%.2f' % x for x in d)}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
enc = eng.encode(specs[0])
torch.cuda.synchronize()
e0.record()
eng.decode_greedy(enc)
e1.record()
torch.cuda.synchronize()
print("alone (incl. cross-K/V GEMMs): %.2f ms" % e0.elapsed_time(e1))
