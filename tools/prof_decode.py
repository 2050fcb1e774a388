#!/usr/bin/env python
"""Small driver for profiling the greedy decoder alone: C2 model, one encode, N decodes.

  python tools/prof_decode.py [--batch 64] [--steps 128] [--reps 3] [--mode cluster] [--workload C2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", default="cluster")
ap.add_argument("--workload", default="C2")
a = ap.parse_args()
os.environ["ASR_B200_DECODE"] = a.mode
from asr_transformer_b200 import workloads as O  # noqa: E402  (workload registry + synthetic inputs)
from asr_transformer_b200.workloads import build_model  # noqa: E402

cfg = O.CONFIGS[a.workload]
dev = torch.device("cuda", 0)
m = build_model(cfg, dev)
spec = O.structured_spectrum(a.batch, cfg.frames, cfg.input_dim, seed=1).to(dev)
eng = m._eng()
enc = eng.encode(spec)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps):
    ev0.record()
    tokens, n, _ = eng.decode_greedy(enc, max_len=a.steps)
    ev1.record()
    torch.cuda.synchronize()
    print(f"decode {a.mode} B={a.batch} L={a.steps}: {ev0.elapsed_time(ev1):.3f} ms "
          f"({1e3 * ev0.elapsed_time(ev1) / a.steps:.1f} us/step)")
print("distinct token rows:", len({tuple(r) for r in tokens.cpu().tolist()}))
