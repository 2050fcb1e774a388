#!/usr/bin/env python
"""Per-phase clock totals of the cluster decoder (asr_decode_profile): python tools/prof_phases.py [batch] [mode] [workload]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_transformer_b200 import workloads as W, lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
os.environ["ASR_B200_DECODE"] = sys.argv[2] if len(sys.argv) > 2 else "cluster2"
cfg = W.CONFIGS[sys.argv[3] if len(sys.argv) > 3 else "C2"]
dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
eng = m._eng()
lib = L.load()
spec = W.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=1).to(dev)
enc = eng.encode(spec)
tokens = torch.empty(B, cfg.decoder_seq_len + 1, dtype=torch.int32, device=dev)
ws = eng._ws(B, 4 * cfg.encoder_seq_len + 3, cfg.decoder_seq_len)
ms, n = (C.c_float * 12)(), (C.c_int32 * 12)()
phase = torch.zeros(148, 16, dtype=torch.int64, device=dev)
for _ in range(2):
    L.check(lib.asr_decode_profile(eng.handle, L.ptr(enc), B, cfg.encoder_seq_len, cfg.decoder_seq_len, L.ptr(ws),
                                   ws.numel(), L.ptr(tokens), ms, n, L.ptr(phase), L.stream()), "profile")
ph = phase.double().cpu()
cp = ph[ph[:, 0] > 0]
Ls = cfg.decoder_seq_len
print(f"{os.environ['ASR_B200_DECODE']} B={B}: {ms[9]:.3f} ms per decode, {len(cp)} CTAs")
print({k: round(float(cp[:, i].mean()) / Ls, 1) for i, k in enumerate(["total", "ring_wait", "exchange_wait", "producer_wait_empty", "stages"])})
names = ["params+cls+argmax", "ln1", "qkv", "self_attn", "attn_finish", "wo", "exchange+ln(x3)", "cross_q", "cross_attn", "w1", "w2"]
print({nm: round(float(cp[:, 5 + i].mean()) / Ls, 1) for i, nm in enumerate(names)})
