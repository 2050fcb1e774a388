#!/usr/bin/env python
"""Stage timeline of one decode step (diagnostic build -DASR_TRACE): consumer warp 0 acquire (begin, end) clocks and the
producer's (wait begin, armed) clocks for step 70 of CTA 0.  python tools/prof_trace.py [batch]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_transformer_b200 import workloads as W, lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
os.environ["ASR_B200_DECODE"] = "cluster"
cfg = W.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = W.build_model(cfg, dev)
eng = m._eng()
lib = L.load()
spec = W.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=1).to(dev)
enc = eng.encode(spec)
tokens = torch.empty(B, cfg.decoder_seq_len + 1, dtype=torch.int32, device=dev)
ws = eng._ws(B, 4 * cfg.encoder_seq_len + 3, cfg.decoder_seq_len)
ms, n = (C.c_float * 12)(), (C.c_int32 * 12)()
phase = torch.zeros(148 * 16 + 1024, dtype=torch.int64, device=dev)
for _ in range(2):
    L.check(lib.asr_decode_profile(eng.handle, L.ptr(enc), B, cfg.encoder_seq_len, cfg.decoder_seq_len, L.ptr(ws),
                                   ws.numel(), L.ptr(tokens), ms, n, L.ptr(phase), L.stream()), "profile")
tr = phase[148 * 16:].cpu().view(2, 256, 2)
cons, prod = tr[0], tr[1]
t0 = int(min(cons[0, 0], prod[0, 0]))
TAGS = ["?", "small", "Wqkv", "selfK", "selfV", "Wo", "Wqc", "crossK", "crossV", "Woc", "W1", "W2", "Wcls", "?", "?", "?"]
print("stage kind | producer: wait_begin armed | consumer: acq_begin acq_end (wait) | lead = acq_begin - armed | to next acquire")
for i in range(256):
    if cons[i, 0] == 0:
        break
    pb, pa, cb, ce = int(prod[i, 0]) - t0, int(prod[i, 1]) - t0, int(cons[i, 0]) - t0, int(cons[i, 1]) - t0
    kind = TAGS[int(prod[i, 0]) & 15]
    nxt = (int(cons[i + 1, 0]) - t0 - cb) if i + 1 < 256 and cons[i + 1, 0] != 0 else 0
    print(f"{i:3d} {kind:7s} | {pb:7d} {pa:7d} | {cb:7d} {ce:7d} ({ce - cb:5d}) | {cb - pa:7d} | {nxt:6d}")
