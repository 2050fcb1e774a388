#!/usr/bin/env python
"""CPU emulation: which bf16 rounding of the decode step flips greedy tokens at argmax near-ties?  The oracle's
KV-cached greedy search (fp32) is re-run with selected tensors rounded to bf16 the way the CUDA path stores them, on the
oracle's own fp32 encoder output.  (Test infrastructure: imports the oracle.)

  python tools/parity_emulation.py [utterances] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from asr_transformer_b200 import workloads as W  # noqa: E402
from oracle import speech_transformer as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = W.CONFIGS["C2"]
torch.set_num_threads(os.cpu_count() or 1)
m = W.build_model(cfg)
sd = W.cpu_state(m)
spec = W.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=11)
enc = O.encode(sd, spec)


def r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def greedy(round_cross_in, round_cross_kv, round_self_kv):
    """O.greedy_kv_cached with rounding hooks (same arithmetic otherwise)."""
    prefix = "decoder"
    Bq, Tp, D = enc.shape
    H = cfg.num_heads
    dh = D // H
    nl = O._num_layers(sd, prefix)
    scale = D ** (-0.5)
    emb, pe, Wc = sd[prefix + "._embedding.weight"], sd[prefix + "._pe.pe"][0], sd[prefix + "._classifier.weight"]

    def heads(t):
        return t.view(Bq, -1, H, dh).transpose(1, 2)

    packs = []
    e_in = r(enc) if round_cross_in else enc
    for l in range(nl):
        lp = f"{prefix}._layers.{l}"
        sq, sk, sv = (O._packed_heads(sd, lp + "._mask_attention", n) for n in ("_q", "_k", "_v"))
        cq, ck, cv = (O._packed_heads(sd, lp + "._cross_attention", n) for n in ("_q", "_k", "_v"))
        k, v = F.linear(e_in, *ck), F.linear(e_in, *cv)
        if round_cross_kv:
            k, v = r(k), r(v)
        packs.append((lp, sq, sk, sv, cq, heads(k), heads(v)))
    tokens = torch.full((Bq, L + 1), cfg.bos_token_id, dtype=torch.int64)
    kc = [torch.zeros(Bq, H, L, dh) for _ in range(nl)]
    vc = [torch.zeros(Bq, H, L, dh) for _ in range(nl)]
    for t in range(L):
        h = emb[tokens[:, t]] + pe[t]
        for l, (lp, sq, sk, sv, cq, ck_x, cv_x) in enumerate(packs):
            a = O.layer_norm(sd, lp + "._norm1", h)
            q = F.linear(a, *sq).view(Bq, H, 1, dh)
            k, v = F.linear(a, *sk).view(Bq, H, dh), F.linear(a, *sv).view(Bq, H, dh)
            kc[l][:, :, t] = r(k) if round_self_kv else k
            vc[l][:, :, t] = r(v) if round_self_kv else v
            kk, vv = kc[l][:, :, :t + 1].clone(), vc[l][:, :, :t + 1].clone()
            kk[:, :, t], vv[:, :, t] = k, v               # the current row is used unrounded (it comes from shared memory)
            s = (q @ kk.transpose(2, 3)) * scale
            o = torch.softmax(s, -1) @ vv
            h = O.linear(sd, lp + "._mask_attention._out_linear", o.transpose(1, 2).reshape(Bq, D)) + h
            a = O.layer_norm(sd, lp + "._norm2", h)
            q = F.linear(a, *cq).view(Bq, H, 1, dh)
            s = (q @ ck_x.transpose(2, 3)) * scale
            o = torch.softmax(s, -1) @ cv_x
            h = O.linear(sd, lp + "._cross_attention._out_linear", o.transpose(1, 2).reshape(Bq, D)) + h
            h = O.feed_forward(sd, lp + "._feedforward", O.layer_norm(sd, lp + "._norm3", h)) + h
        tokens[:, t + 1] = F.linear(h, Wc).argmax(-1)
    return tokens


ref = greedy(False, False, False)
for name, args in (("cross K/V inputs (enc_out) bf16", (True, False, False)),
                   ("cross K/V stored bf16", (False, True, False)),
                   ("self K/V cache bf16", (False, False, True)),
                   ("all three (the CUDA path's storage)", (True, True, True))):
    t = greedy(*args)
    same = int((t == ref).all(-1).sum())
    print(f"{name:<40s} identical {same}/{B}")
