#!/usr/bin/env python
"""Bring-up check of the tcgen05 cluster decoder against the per-kernel (graph) step: tokens + step logits.
  python tools/check_cluster2.py [workload] [batch] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_transformer_b200 import workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "T0"
cfg = W.CONFIGS[name]
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg.batch
L = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.decoder_seq_len
mode2 = sys.argv[4] if len(sys.argv) > 4 else "cluster2"
dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
spec = W.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=23).to(dev)
os.environ["ASR_B200_DECODE"] = "graph"
tg, ng, lg = m.greedy_decode(spec, max_len=L, return_logits=True)
torch.cuda.synchronize()
os.environ["ASR_B200_DECODE"] = mode2
tc, nc, lc = m.greedy_decode(spec, max_len=L, return_logits=True)
torch.cuda.synchronize()
same = (tg == tc).all(-1)
d = (lg - lc).abs()
first = [(int(b), int((tg[b] != tc[b]).nonzero()[0])) for b in range(B) if not same[b]]
print(f"{name} B={B} L={L} {mode2} vs graph: identical {int(same.sum())}/{B}; step logits max |d| {d.max():.3e} mean {d.mean():.3e}; "
      f"finite {bool(torch.isfinite(lc).all())}; first divergences {first[:8]}")
if int(same.sum()) < B:
    b, pos = first[0]
    print("  step-0 logits diff:", float((lg[b, 0] - lc[b, 0]).abs().max()), " step", pos - 1, "diff:",
          float((lg[b, pos - 1] - lc[b, pos - 1]).abs().max()))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng = m._eng()
enc = eng.encode(spec)
for i in range(3):
    ev0.record()
    eng.decode_greedy(enc, max_len=L)
    ev1.record()
    torch.cuda.synchronize()
print(f"  decode {mode2}: {ev0.elapsed_time(ev1):.3f} ms ({1e3 * ev0.elapsed_time(ev1) / L:.1f} us/step)")
