#!/usr/bin/env python
"""C5 (d_model 512, clusters of 8 CTAs): one decode launch of 64 mixed-length utterances; prints the launch time and
a hash of the tokens, so that runs with ASR_B200_CLUSTER_GU=4 / 8 can be compared."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_transformer_b200 import workloads as W  # noqa: E402

dev = torch.device("cuda", 0)
c5 = W.CONFIGS["C5"]
m5 = W.build_model(c5, dev)
g = torch.Generator().manual_seed(5)
lens = torch.randint(400, 1001, (64,), generator=g)
x = W.structured_spectrum(64, 1000, c5.input_dim, seed=500, lengths=lens).to(dev)
ln = lens.to(dev)
for _ in range(2):
    tok, n = m5.greedy_decode(x, lengths=ln)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tok, n = m5.greedy_decode(x, lengths=ln)
e1.record()
torch.cuda.synchronize()
t = tok.cpu().to(torch.int32).contiguous()
print(f"GU={os.environ.get('ASR_B200_CLUSTER_GU', 'auto')}: {e0.elapsed_time(e1):.2f} ms per 64 utterances, tokens sha "
      f"{hashlib.sha256(t.numpy().tobytes()).hexdigest()[:16]}, distinct rows {len({tuple(r) for r in t.tolist()})}")
torch.save(t, os.path.join(ROOT, "gpurun_out", f"c5_tokens_gu{os.environ.get('ASR_B200_CLUSTER_GU', 'auto')}.pt"))
