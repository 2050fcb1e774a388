#!/usr/bin/env python
"""Timeline of the pipelined serving loop (Transformer.greedy_decode_batches, ASR_B200_PIPE_TRACE=1): when every group's
front-end + encoder, cross-K/V preparation and decode launch start and end.  python tools/prof_pipeline.py [batches]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ASR_B200_PIPE_TRACE"] = "1"
import torch  # noqa: E402
from asr_transformer_b200 import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg = W.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
specs = [W.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=100 + i).to(dev) for i in range(8)]
for rep in range(2):   # the second pass is the warm one
    if rep:
        print("---- warm pass", file=sys.stderr)
    for _ in m.greedy_decode_batches((specs[i % 8] for i in range(n)), to_host=False):
        pass
torch.cuda.synchronize()
