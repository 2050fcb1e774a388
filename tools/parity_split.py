#!/usr/bin/env python
"""Where do the near-tie token flips come from?  C2, B utterances: (a) whole path on the GPU, (b) GPU decoder fed with the
oracle's fp32 encoder output, both against the oracle's tokens.  (Test infrastructure: imports the oracle.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asr_transformer_b200 import workloads as W  # noqa: E402
from oracle import speech_transformer as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = W.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
sd = W.cpu_state(m)
spec = W.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=11)
torch.set_num_threads(os.cpu_count() or 1)
enc_ref = O.encode(sd, spec)
tok_ref, lg_ref = O.greedy_kv_cached(sd, enc_ref, cfg)
eng = m._eng()
enc = m.encode(spec.to(dev))
print("enc_out max|d| %.3e mean|d| %.3e" % ((enc.cpu() - enc_ref).abs().max(), (enc.cpu() - enc_ref).abs().mean()))
for name, e in (("gpu encoder + gpu decoder", enc), ("oracle fp32 enc_out + gpu decoder", enc_ref.to(dev))):
    tok, _, lg = eng.decode_greedy(e, want_logits=True)
    r = O.compare_tokens(tok_ref, lg_ref, tok, 2e-2)
    same = [b for b in range(B) if torch.equal(tok_ref[b], tok[b].cpu().long())]
    d = (lg[same].cpu() - lg_ref[same]).abs()
    print(f"{name}: identical {r['identical']}/{B}, hard {len(r['hard'])}, max near-tie margin "
          f"{max([x[2] for x in r['near_tie']] or [0]):.2e}; logits max|d| {d.max():.2e} mean|d| {d.mean():.2e}")
