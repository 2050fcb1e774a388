#!/usr/bin/env python
"""How much does each encoder-side kernel slow the cluster decoder down when it runs beside it (on the 20 SMs the
decoder leaves free)?  The decode launch of the bench (128 utterances) is timed alone and under a looping background
operator on a side stream."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asr_transformer_b200 import lib as L  # noqa: E402
from asr_transformer_b200 import workloads as O  # noqa: E402  (workload registry + synthetic inputs)
from asr_transformer_b200.workloads import build_model  # noqa: E402

cfg = O.CONFIGS["C2"]
dev = torch.device("cuda", 0)
m = build_model(cfg, dev)
lib = L.load()
B = 128
spec = O.structured_spectrum(B, cfg.frames, cfg.input_dim, seed=1).to(dev)
eng = m._eng()
enc = eng.encode(spec)
side = torch.cuda.Stream(dev)
Tp, D, M = 249, 256, B * 249


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def gemm_fn(N, K, relu=0, res=False, f32=False, b16=True):
    x = rnd(M, K, dtype=torch.bfloat16)
    w = rnd(N, K, dtype=torch.bfloat16, scale=K ** -0.5)
    bias = rnd(N)
    r = rnd(M, N) if res else None
    y32 = torch.empty(M, N, device=dev) if f32 else None
    y16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if b16 else None
    return lambda: L.check(lib.asr_gemm_bf16(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(r), None, Tp, M, N, K, relu,
                                             L.ptr(y32), L.ptr(y16), 0, L.stream()), "gemm")


qkv = rnd(B, Tp, 3 * D, dtype=torch.bfloat16)
aout = torch.empty(B, Tp, D, dtype=torch.bfloat16, device=dev)


def attn_fn():
    base = qkv.data_ptr()
    L.check(lib.asr_attention(C.c_void_p(base), 3 * D, Tp * 3 * D, C.c_void_p(base + 2 * D), 3 * D, Tp * 3 * D,
                              C.c_void_p(base + 4 * D), 3 * D, Tp * 3 * D, L.ptr(aout), D, Tp * D, B, 4, Tp, Tp,
                              D ** -0.5, 0, None, None, None, None, 1, 0, L.stream()), "attention")


xl = rnd(M, D)
g, bt = rnd(D), rnd(D)
y16 = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
big_a = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
big_b = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
bg = {
    "nothing": None,
    "full encoder": lambda: eng.encode(spec, ws_tag="bg"),
    "d2d copy 64 MB": lambda: big_b.copy_(big_a),
    "gemm ffn1 (N=1024,K=256)": gemm_fn(1024, 256, relu=1),
    "gemm ffn2+res (N=256,K=1024)": gemm_fn(256, 1024, res=True, f32=True, b16=False),
    "attention": attn_fn,
    "layernorm": lambda: L.check(lib.asr_layernorm(L.ptr(xl), L.ptr(g), L.ptr(bt), M, D, None, L.ptr(y16), L.stream()), "ln"),
    "conv front-end": lambda: m.input_layer(spec[:64]),
}
for name, fn in bg.items():
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode_greedy(enc)
        e1.record()
        n = 0
        if fn is not None:
            with torch.cuda.stream(side):
                while not e1.query() and n < 4000:
                    fn()
                    n += 1
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:<32s} decode {min(ts):7.3f} ms (background launches {n})")
