#!/usr/bin/env python
"""Small driver for profiling the encoder alone: C2 model, N encodes (conv front-end + 6 encoder layers)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--workload", default="C2")
a = ap.parse_args()
from asr_transformer_b200 import workloads as O  # noqa: E402  (workload registry + synthetic inputs)
from asr_transformer_b200.workloads import build_model  # noqa: E402

cfg = O.CONFIGS[a.workload]
dev = torch.device("cuda", 0)
m = build_model(cfg, dev)
spec = O.structured_spectrum(a.batch, cfg.frames, cfg.input_dim, seed=1).to(dev)
eng = m._eng()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(a.reps):
    ev0.record()
    enc = eng.encode(spec)
    ev1.record()
    torch.cuda.synchronize()
    print(f"encode B={a.batch}: {ev0.elapsed_time(ev1):.3f} ms")
