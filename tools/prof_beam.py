#!/usr/bin/env python
"""Beam search timing at the C2 workload: B utterances x beam hypotheses, L steps (eager per-kernel step path)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asr_transformer_b200 import workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--beam", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg = W.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
spec = W.structured_spectrum(a.batch, cfg.frames, cfg.input_dim, seed=1).to(dev)
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tok, sc = m.beam_search(spec, beam=a.beam)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"beam search B={a.batch} beam={a.beam} L={cfg.decoder_seq_len}: {ms:.1f} ms = {1e3 * a.batch / ms:.0f} utt/s")
tg, _ = m.greedy_decode(spec, stop_at_eos=True)
print("best hypothesis == greedy for", int((tok[:, 0] == tg).all(-1).sum()), "of", a.batch, "utterances; mean score gain",
      float((sc[:, 0] - m.beam_search(spec, beam=1)[1][:, 0]).mean()))
