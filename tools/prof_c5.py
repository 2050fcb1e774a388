#!/usr/bin/env python
"""C5 mixed-length serving: length buckets of 16 (each padded to its own longest utterance, one decode launch per bucket)
against buckets of 32 / one padded batch of 64 (more padding in the encoder, fewer decode launches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_transformer_b200 import workloads as W  # noqa: E402
from asr_transformer_b200.parallel import bucket_by_length  # noqa: E402

dev = torch.device("cuda", 0)
c5 = W.CONFIGS["C5"]
m5 = W.build_model(c5, dev)
g = torch.Generator().manual_seed(5)
lens = torch.randint(400, 1001, (64,), generator=g)
mine = list(range(64))
for bs in (16, 32, 64):
    items = []
    for b in bucket_by_length([int(lens[i]) for i in mine], bs):
        idx = [mine[j] for j in b]
        ln = lens[idx]
        T = int(ln.max())
        x = W.structured_spectrum(len(idx), T, c5.input_dim, seed=500 + idx[0], lengths=ln).to(dev)
        items.append((x, ln.to(dev)))
    for _ in range(2):
        for x, ln in items:
            m5.greedy_decode(x, lengths=ln)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for x, ln in items:
            m5.greedy_decode(x, lengths=ln)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    eng = m5._eng()
    t_enc = 0.0
    for x, ln in items:
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.encode(x, lengths=ln) if "lengths" in eng.encode.__code__.co_varnames else eng.encode(x)
        b_.record()
        torch.cuda.synchronize()
        t_enc += a.elapsed_time(b_)
    print(f"bucket size {bs}: {len(items)} launches, {ms:.2f} ms per 64 utterances = {64 / ms * 1e3:.0f} utt/s (encode part {t_enc:.2f} ms)")
