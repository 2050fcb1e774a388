#!/usr/bin/env python
"""Times the encoder self-attention op alone (C2 shape: T' = 250, 4 heads) through the C ABI: us per launch."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_transformer_b200 import lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="64,256,512")
ap.add_argument("--S", type=int, default=250)
ap.add_argument("--H", type=int, default=4)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = L.load()
H, S = a.H, a.S
for B in [int(x) for x in a.batches.split(",")]:
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(B, S, 3 * H * 64, generator=g) * 1.0).to(torch.float16).to(dev)
    out = torch.zeros(B, S, 2 * H * 64, dtype=torch.float16, device=dev)
    scale = (64 * H) ** -0.5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run():
        L.check(lib.asr_attention(qkv.data_ptr(), 3 * H * 64, S * 3 * H * 64, qkv.data_ptr() + 2 * H * 64, 3 * H * 64,
                                  S * 3 * H * 64, qkv.data_ptr() + 4 * H * 64, 3 * H * 64, S * 3 * H * 64, out.data_ptr(),
                                  2 * H * 64, S * 2 * H * 64, B, H, S, S, scale, 0, None, None, None, None, 1, 0,
                                  L.stream()), "attention")
    for _ in range(3):
        run()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    flop = 4.0 * S * S * 64 * B * H
    print(f"attention B={B} S={S} H={H}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  "
          f"{flop / ts[len(ts)//2] / 1e6:.0f} TF/s")
