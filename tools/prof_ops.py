#!/usr/bin/env python
"""Warm, back-to-back timing of the encoder-side operators at the C2 shapes (B utterances of 10 s), each through its
C-ABI entry point: n launches between two CUDA events on the launching stream.

  python tools/prof_ops.py [--batch 64] [--reps 30]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from asr_transformer_b200 import lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = L.load()
B, Tp, D, FF, H = a.batch, 249, 256, 1024, 4
M = B * Tp


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def timeit(name, fn, flops=0.0, bytes_=0.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / a.reps
    print(f"{name:<44s} {us:8.2f} us  {flops / us / 1e6:8.1f} TF/s  {bytes_ / us / 1e3:8.1f} GB/s")
    return us


def gemm(N, K, relu=0, res=False, pe=False, f32=False, b16=True, name=""):
    x = rnd(M, K, dtype=torch.float16)
    w = rnd((N + 63) // 64 * 64, K, dtype=torch.float16, scale=K ** -0.5)
    bias = rnd(N)
    r = rnd(M, N) if res else None
    p = rnd(Tp, N) if pe else None
    y32 = torch.empty(M, N, device=dev) if f32 else None
    y16 = torch.empty(M, N, dtype=torch.float16, device=dev) if b16 else None
    by = M * K * 2 + N * K * 2 + (M * N * 4 if res else 0) + (M * N * 4 if f32 else 0) + (M * N * 2 if b16 else 0)

    def fn():
        L.check(lib.asr_gemm_f16(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(r), L.ptr(p), Tp, M, N, K, relu, L.ptr(y32),
                                  L.ptr(y16), 0, L.stream()), "gemm")
    return timeit(f"gemm {name} M={M} N={N} K={K}", fn, 2.0 * M * N * K, by)


tot = 0.0
t = gemm(256, 1216, pe=True, f32=True, b16=False, name="lin_in+pe")
tot += t
t_qkv = gemm(768, 256, name="qkv")
t_out = gemm(256, 256, res=True, f32=True, b16=False, name="out+res")
t_f1 = gemm(1024, 256, relu=1, name="ffn1+relu")
t_f2 = gemm(256, 1024, res=True, f32=True, b16=False, name="ffn2+res")
t_ckv = gemm(512, 256, name="cross kv")

qkv = rnd(B, Tp, 3 * D, dtype=torch.float16)
out = torch.empty(B, Tp, D, dtype=torch.float16, device=dev)


def attn():
    q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
    L.check(lib.asr_attention(L.ptr(q), 3 * D, Tp * 3 * D, L.ptr(k), 3 * D, Tp * 3 * D, L.ptr(v), 3 * D, Tp * 3 * D,
                              L.ptr(out), D, Tp * D, B, H, Tp, Tp, D ** -0.5, 0, None, None, None, None, 1, 0,
                              L.stream()), "attention")


# strided views: pass raw pointers with offsets
def attn_ptr():
    base = qkv.data_ptr()
    import ctypes as C
    L.check(lib.asr_attention(C.c_void_p(base), 3 * D, Tp * 3 * D, C.c_void_p(base + 2 * D), 3 * D, Tp * 3 * D,
                              C.c_void_p(base + 4 * D), 3 * D, Tp * 3 * D, L.ptr(out), D, Tp * D, B, H, Tp, Tp,
                              D ** -0.5, 0, None, None, None, None, 1, 0, L.stream()), "attention")


t_attn = timeit(f"attention B={B} H={H} S={Tp}", attn_ptr, 4.0 * B * H * Tp * Tp * 64, B * Tp * 4 * D * 2)
x = rnd(M, D)
g, bt = rnd(D), rnd(D)
y16 = torch.empty(M, D, dtype=torch.float16, device=dev)
t_ln = timeit("layernorm -> bf16", lambda: L.check(lib.asr_layernorm(L.ptr(x), L.ptr(g), L.ptr(bt), M, D, None,
                                                                      L.ptr(y16), L.stream()), "ln"), 0, M * D * 6)

# conv front-end
F_, T_ = 80, 1000
spec = rnd(B, 1, F_, T_)
w1, b1 = rnd(9, 64), rnd(64)
w2 = rnd(36, 8, 32, 4, dtype=torch.float16)
b2 = rnd(64)
nws = lib.asr_conv_workspace_bytes(B, F_, T_)
ws = torch.empty(nws, dtype=torch.uint8, device=dev)
z = torch.empty(B, Tp, 19 * 64, dtype=torch.float16, device=dev)
t_conv = timeit("conv front-end (conv1 + conv2)",
                lambda: L.check(lib.asr_conv_frontend(L.ptr(spec), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), B, F_, T_,
                                                      L.ptr(ws), nws, L.ptr(z), L.stream()), "conv"),
                2.0 * B * (9 * 64 * 39 * 499 + 9 * 64 * 64 * 19 * 249), B * (F_ * T_ * 4 + 2 * 39 * 499 * 64 * 2 + Tp * 1216 * 2))
layer = t_ln * 2 + t_qkv + t_attn + t_out + t_f1 + t_f2
print(f"sum: conv {t_conv:.0f} + lin_in {t:.0f} + 6 x layer {layer:.0f} + ln {t_ln:.0f} = "
      f"{t_conv + t + 6 * layer + t_ln:.0f} us;  cross kv 6 x {t_ckv:.0f}")
