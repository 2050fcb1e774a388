"""Operator-level parity on the GPU: every kernel through the C-ABI against a plain fp32 reference of the same op
(torch on the device for generic ops, the CPU oracle / golden fixtures for reference-specific behaviour)."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from asr_transformer_b200 import engine as E
from asr_transformer_b200 import lib as L
from oracle import speech_transformer as O
from tests.util import TOL_FP32, assert_close, golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def lib():
    return L.load()


def sync():
    torch.cuda.synchronize()


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ----------------------------------------------------------------------------------------------- tcgen05 probe
@pytest.mark.parametrize("N", [32, 64, 128])
def test_umma_probe_k_major(N):
    a = rnd(128, 64, seed=1, dtype=torch.float16)
    b = rnd(N, 64, seed=2, dtype=torch.float16)
    d = torch.zeros(128, N, device=DEV)
    L.check(lib().asr_umma_probe(L.ptr(a), L.ptr(b), L.ptr(d), N, 0, L.stream()))
    sync()
    assert_close(d, a.float() @ b.float().t(), 1e-3, 1e-4, "umma K-major")


def test_umma_probe_mn_major():
    a = rnd(128, 64, seed=3, dtype=torch.float16)
    b = rnd(64, 64, seed=4, dtype=torch.float16)       # [k, n], n contiguous (a V tile)
    d = torch.zeros(128, 64, device=DEV)
    L.check(lib().asr_umma_probe(L.ptr(a), L.ptr(b), L.ptr(d), 64, 1, L.stream()))
    sync()
    assert_close(d, a.float() @ b.float(), 1e-3, 1e-4, "umma MN-major B")


# ----------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,D", [(1, 128), (37, 256), (1000, 256), (15936, 256), (513, 512), (9, 1024)])
def test_layernorm(rows, D):
    x = rnd(rows, D, seed=5, scale=3.0) + 0.7
    g, b = rnd(D, seed=6) * 0.2 + 1.0, rnd(D, seed=7) * 0.1
    y32 = torch.empty_like(x)
    y16 = torch.empty(rows, D, dtype=torch.float16, device=DEV)
    L.check(lib().asr_layernorm(L.ptr(x), L.ptr(g), L.ptr(b), rows, D, L.ptr(y32), L.ptr(y16), L.stream()))
    sync()
    ref = F.layer_norm(x, (D,), g, b, 1e-5)
    assert_close(y32, ref, TOL_FP32, TOL_FP32, "layernorm fp32")
    assert torch.equal(y16, y32.to(torch.float16))


def test_layernorm_rejects_bad_width():
    x = rnd(4, 96)
    rc = lib().asr_layernorm(L.ptr(x), L.ptr(x), L.ptr(x), 4, 96, L.ptr(x), None, L.stream())
    assert rc == -2 and b"multiple of 128" in lib().asr_last_error()


# ----------------------------------------------------------------------------------------------- GEMM
def run_gemm(x, w, bias, res, pe, period, relu, impl, N):
    M, K = x.shape
    y32 = torch.zeros(M, N, device=DEV)
    y16 = torch.zeros(M, N, dtype=torch.float16, device=DEV)
    L.check(lib().asr_gemm_f16(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(res), L.ptr(pe), period, M, N, K, relu,
                                L.ptr(y32), L.ptr(y16), impl, L.stream()), "gemm")
    sync()
    return y32, y16


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 256, 256), (1000, 768, 256), (129, 1024, 256),
                                   (777, 256, 1024), (2000, 256, 1216), (257, 250, 128), (15936, 768, 256),
                                   (5, 512, 512)])
def test_gemm_tc_plain(M, N, K):
    x = rnd(M, K, seed=8, dtype=torch.float16)
    npad = (N + 63) // 64 * 64
    w = torch.zeros(npad, K, dtype=torch.float16, device=DEV)
    w[:N] = rnd(N, K, seed=9, scale=K ** -0.5, dtype=torch.float16)
    y32, y16 = run_gemm(x, w, None, None, None, 1, 0, 0, N)
    ref = x.float() @ w[:N].float().t()
    assert_close(y32, ref, 2e-3, 2e-4, f"gemm {M}x{N}x{K}")
    assert_close(y16, ref, 3e-2, 4e-3, "gemm bf16 out")
    n32, _ = run_gemm(x, w, None, None, None, 1, 0, 1, N)
    assert_close(n32, ref, 2e-3, 2e-4, "naive gemm")


@pytest.mark.parametrize("relu,use_res,use_pe", [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)])
def test_gemm_tc_epilogues(relu, use_res, use_pe):
    M, N, K, period = 498, 256, 256, 249
    x = rnd(M, K, seed=10, dtype=torch.float16)
    w = rnd(N, K, seed=11, scale=K ** -0.5, dtype=torch.float16)
    bias = rnd(N, seed=12)
    res = rnd(M, N, seed=13) if use_res else None
    pe = rnd(period, N, seed=14) if use_pe else None
    y32, y16 = run_gemm(x, w, bias, res, pe, period, relu, 0, N)
    ref = x.float() @ w.float().t() + bias
    if relu:
        ref = ref.relu()
    if use_pe:
        ref = ref + pe.repeat(M // period, 1)
    if use_res:
        ref = ref + res
    assert_close(y32, ref, 2e-3, 2e-4, "gemm epilogue")
    assert torch.equal(y16, y32.to(torch.float16))


@pytest.mark.parametrize("M,N,K,relu", [(300, 256, 256, 0), (1000, 1024, 256, 1), (515, 256, 1024, 0), (77, 250, 1216, 0)])
def test_gemm_split_is_fp32_accurate(M, N, K, relu):
    """hi | lo split activations (the model path's operand format): the tensor-core product equals the fp64 product of
    the fp32 activations to fp32 rounding, and the hi | lo output pair reassembles the fp32 result to 2^-21."""
    x = rnd(M, K, seed=15, scale=2.0) + 0.3
    npad = (N + 63) // 64 * 64
    w = torch.zeros(npad, K, dtype=torch.float16, device=DEV)
    w[:N] = rnd(N, K, seed=16, scale=K ** -0.5, dtype=torch.float16)
    bias = rnd(N, seed=17)
    y32 = torch.zeros(M, N, device=DEV)
    y16 = torch.zeros(M, 2 * N, dtype=torch.float16, device=DEV) if N % 8 == 0 else None
    ws = torch.empty(M * K * 4, dtype=torch.uint8, device=DEV)
    L.check(lib().asr_gemm_split(L.ptr(x), L.ptr(w), L.ptr(bias), M, N, K, relu, L.ptr(y32), L.ptr(y16), L.ptr(ws),
                                 ws.numel(), L.stream()), "gemm_split")
    sync()
    ref = x.double() @ w[:N].double().t() + bias.double()
    if relu:
        ref = ref.relu()
    assert_close(y32, ref.float(), 2e-5 * K ** 0.5, 2e-6 * K ** 0.5, "split gemm (fp32-accurate)")
    if y16 is not None:
        back = y16[:, :N].float() + y16[:, N:].float()
        assert_close(back, y32, 4e-6, 4e-7, "hi | lo output pair")


@pytest.mark.parametrize("M,K,use_res,use_pe", [(300, 256, 1, 0), (1000, 1024, 1, 0), (515, 1216, 0, 1), (15936, 256, 1, 0),
                                                (77, 64, 0, 0)])
def test_gemm_ln_fused(M, K, use_res, use_pe):
    """Full-row GEMM + fused LayerNorm epilogue (the replacement of the LayerNorm launches) against fp64 torch."""
    N, period = 256, 249
    x = rnd(M, K, seed=18, scale=1.5) + 0.2
    w = rnd(N, K, seed=19, scale=K ** -0.5, dtype=torch.float16)
    bias = rnd(N, seed=20)
    res = (rnd(M, N, seed=21, scale=2.0) + 1.0) if use_res else None
    pe = rnd(period, N, seed=22) if use_pe else None
    g, b = rnd(N, seed=23) * 0.2 + 1.0, rnd(N, seed=24) * 0.1
    h_out = torch.zeros(M, N, device=DEV)
    y32 = torch.zeros(M, N, device=DEV)
    y16 = torch.zeros(M, 2 * N, dtype=torch.float16, device=DEV)
    ws = torch.empty(M * K * 4, dtype=torch.uint8, device=DEV)
    L.check(lib().asr_gemm_ln(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(res), L.ptr(pe), period, L.ptr(g), L.ptr(b), M, N, K,
                              L.ptr(h_out), L.ptr(y32), L.ptr(y16), L.ptr(ws), ws.numel(), L.stream()), "gemm_ln")
    sync()
    h = x.double() @ w.double().t() + bias.double()
    if use_pe:
        h = h + pe.double().repeat((M + period - 1) // period, 1)[:M]
    if use_res:
        h = h + res.double()
    ref = F.layer_norm(h, (N,), g.double(), b.double(), 1e-5)
    assert_close(h_out, h.float(), 2e-5 * K ** 0.5, 2e-6 * K ** 0.5, "fused gemm: residual row")
    assert_close(y32, ref.float(), 3e-5 * K ** 0.5, 3e-6 * K ** 0.5, "fused LayerNorm fp32")
    assert_close(y16[:, :N].float() + y16[:, N:].float(), y32, 4e-6, 4e-7, "fused LayerNorm hi | lo")


# ----------------------------------------------------------------------------------------------- attention core
def attn_reference(q, k, v, scale, causal=False, k_lens=None, q_valid=None, k_valid=None, dense=None):
    """q (B,Sq,H,64) etc. fp32 math on the bf16 values; masked rows -> zeros (reference layers.py:20-27)."""
    B, Sq, H, _ = q.shape
    Sk = k.shape[1]
    s = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    m = torch.zeros(B, 1, Sq, Sk, dtype=torch.bool, device=q.device)
    if causal:
        m = m | torch.triu(torch.ones(Sq, Sk, dtype=torch.bool, device=q.device), 1)
    if k_lens is not None:
        m = m | (torch.arange(Sk, device=q.device)[None, None, None, :] >= k_lens[:, None, None, None])
    if k_valid is not None:
        m = m | (k_valid[:, None, None, :] == 0)
    if q_valid is not None:
        m = m | (q_valid[:, None, :, None] == 0)
    if dense is not None:
        m = m | (dense[:, None] != 0 if dense.shape[0] == B else dense[None] != 0)
    s = s.masked_fill(m, float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, -1))
    return torch.einsum("bhqk,bkhd->bqhd", p, v.float())


def run_attn(q, k, v, scale, impl=0, causal=0, k_lens=None, q_valid=None, k_valid=None, dense=None):
    B, Sq, H, _ = q.shape
    Sk = k.shape[1]
    out = torch.full((B, Sq, H * 64), float("nan"), dtype=torch.float16, device=DEV)
    mask_b = 1 if dense is None else dense.shape[0]
    L.check(lib().asr_attention(L.ptr(q), H * 64, Sq * H * 64, L.ptr(k), H * 64, Sk * H * 64, L.ptr(v), H * 64,
                                Sk * H * 64, L.ptr(out), H * 64, Sq * H * 64, B, H, Sq, Sk, scale, causal,
                                L.ptr(k_lens), L.ptr(q_valid), L.ptr(k_valid), L.ptr(dense), mask_b, impl, L.stream()),
            "attention")
    sync()
    return out.view(B, Sq, H, 64)


@pytest.mark.parametrize("B,H,Sq,Sk", [(1, 1, 128, 128), (2, 2, 49, 49), (3, 4, 249, 249), (2, 4, 16, 249),
                                       (1, 8, 749, 749), (2, 2, 130, 1)])
@pytest.mark.parametrize("impl", [0, 1])
def test_attention_nomask(B, H, Sq, Sk, impl):
    q, k, v = (rnd(B, S, H, 64, seed=20 + i, dtype=torch.float16) for i, S in enumerate((Sq, Sk, Sk)))
    scale = (64 * H) ** -0.5          # emb_dim ** -0.5, reference layers.py:20
    out = run_attn(q, k, v, scale, impl)
    assert_close(out, attn_reference(q, k, v, scale), 2e-2, 2e-3, f"attention impl={impl}")


@pytest.mark.parametrize("B,H,Sq,Sk", [(40, 4, 250, 250), (37, 4, 128, 250), (9, 8, 300, 256), (150, 1, 77, 33)])
def test_attention_persistent_many_items(B, H, Sq, Sk):
    """Sk <= 256 without causal / byte masks runs the persistent ping-pong kernel (attn_ts_kernel): every CTA walks
    several (utterance, head, q tile) items, K/V are reloaded when the head changes, ragged key lengths incl. 0
    (reference layers.py:20-27 with the key-padding mask of dataset.py:53-55)."""
    q, k, v = (rnd(B, S, H, 64, seed=40 + i, scale=1.5, dtype=torch.float16) for i, S in enumerate((Sq, Sk, Sk)))
    scale = (64 * H) ** -0.5
    out = run_attn(q, k, v, scale)
    assert_close(out, attn_reference(q, k, v, scale), 2e-2, 2e-3, "persistent attention")
    g = torch.Generator().manual_seed(B)
    k_lens = torch.randint(0, Sk + 1, (B,), generator=g, dtype=torch.int32)
    k_lens[0], k_lens[-1] = Sk, 0
    k_lens = k_lens.to(DEV)
    out = run_attn(q, k, v, scale, k_lens=k_lens)
    assert_close(out, attn_reference(q, k, v, scale, k_lens=k_lens), 2e-2, 2e-3, "persistent attention, k_lens")
    assert (out[-1] == 0).all()
    assert torch.equal(out, run_attn(q, k, v, scale, k_lens=k_lens))          # run-to-run identical


@pytest.mark.parametrize("impl", [0, 1])
def test_attention_masks(impl):
    B, H, S = 3, 2, 200
    q, k, v = (rnd(B, S, H, 64, seed=30 + i, scale=2.0, dtype=torch.float16) for i in range(3))
    scale = 128 ** -0.5
    out = run_attn(q, k, v, scale, impl, causal=1)
    assert_close(out, attn_reference(q, k, v, scale, causal=True), 2e-2, 2e-3, "causal")
    k_lens = torch.tensor([200, 77, 0], dtype=torch.int32, device=DEV)
    out = run_attn(q, k, v, scale, impl, k_lens=k_lens)
    ref = attn_reference(q, k, v, scale, k_lens=k_lens)
    assert_close(out, ref, 2e-2, 2e-3, "k_lens")
    assert (out[2] == 0).all()                       # no valid key -> zeros, not NaN
    valid = (torch.rand(B, S, generator=torch.Generator().manual_seed(5)) > 0.3).to(torch.uint8).to(DEV)
    out = run_attn(q, k, v, scale, impl, causal=1, q_valid=valid, k_valid=valid)
    assert_close(out, attn_reference(q, k, v, scale, causal=True, q_valid=valid, k_valid=valid), 2e-2, 2e-3,
                 "decoder-forward style mask")
    dense = (torch.rand(B, S, S, generator=torch.Generator().manual_seed(6)) > 0.5).to(torch.uint8)
    dense[:, 7] = 1
    dense = dense.to(DEV)
    out = run_attn(q, k, v, scale, impl, dense=dense)
    assert_close(out, attn_reference(q, k, v, scale, dense=dense), 2e-2, 2e-3, "dense mask")
    assert (out[:, 7] == 0).all()
    out = run_attn(q, k, v, scale, impl, dense=dense[:1].contiguous())
    assert_close(out, attn_reference(q, k, v, scale, dense=dense[:1]), 2e-2, 2e-3, "broadcast dense mask")


# ----------------------------------------------------------------------------------------------- MHA / FFN modules
def test_mha_module_against_reference_fixture():
    import asr_transformer_b200 as A
    fx = golden("ops_mha.pt")
    mha = A.MHA(fx["H"], fx["D"], 0.1)
    mha.load_state_dict(fx["state"], strict=True)
    mha = mha.to(DEV).eval()
    x, src = fx["x"].to(DEV), fx["src"].to(DEV)
    assert_close(mha(x), fx["self_nomask"], 2e-2, 3e-3, "self")
    assert_close(mha(x, src), fx["cross_nomask"], 2e-2, 3e-3, "cross")
    assert_close(mha(x, src, fx["keypad"].to(DEV)), fx["cross_keypad"], 2e-2, 3e-3, "key padding")
    assert_close(mha(x, attention_mask=fx["causal"].to(DEV)), fx["self_causal"], 2e-2, 3e-3, "2-D uint8 causal")
    out = mha(x, attention_mask=fx["full_rows"].to(DEV))
    assert_close(out, fx["self_full_rows"], 2e-2, 3e-3, "fully masked rows")
    bias = fx["state"]["_out_linear.bias"].to(DEV)
    assert_close(out[:, 5], bias.expand(3, -1), 1e-6, 1e-6, "masked row == out-projection bias")
    # single head drop-in
    head = mha._heads[1]
    sd = {"m." + k: v for k, v in fx["state"].items()}
    hp = "m._heads.1"
    q = O.linear(sd, hp + "._q", fx["x"])
    k = O.linear(sd, hp + "._k", fx["x"])
    v = O.linear(sd, hp + "._v", fx["x"])
    ref = torch.softmax(q.bmm(k.transpose(1, 2)) * fx["D"] ** -0.5, -1).bmm(v)
    assert_close(head(x), ref, 2e-2, 3e-3, "MHAHead")


def test_ffn_and_layernorm_modules():
    import asr_transformer_b200 as A
    fx = golden("ops_mha.pt")
    ff = A.FeedForward(fx["D"], 256, 0.1)
    ff.load_state_dict(fx["ffn_state"], strict=True)
    ff = ff.to(DEV).eval()
    assert_close(ff(fx["x"].to(DEV)), fx["ffn_out"], 2e-2, 3e-3, "FeedForward")
    ln = A.LayerNorm(fx["D"]).to(DEV).eval()
    x = fx["x"].to(DEV) * 3 + 1
    assert_close(ln(x), F.layer_norm(x, (fx["D"],)), TOL_FP32, TOL_FP32, "LayerNorm module")


@pytest.mark.parametrize("rows,FF", [(300, 1024), (15936, 1024), (77, 256), (1000, 2048)])
def test_ffn_fused_matches_fp64_and_two_gemm_path(rows, FF, monkeypatch):
    """The fused FFN kernel (hidden activation kept on the SM) against fp64 torch and against the two-GEMM path."""
    import asr_transformer_b200 as A
    D = 256
    torch.manual_seed(5)
    ff = A.FeedForward(D, FF, 0.1).eval()
    with torch.no_grad():
        for p_ in ff.parameters():
            O.bf16_representable_(p_)
    ff = ff.to(DEV)
    x = rnd(rows, D, seed=70, scale=1.5) + 0.1
    out = ff(x)
    sync()
    monkeypatch.setenv("ASR_B200_FUSE_FFN", "0")
    ref2 = ff(x)      # (the switch is read once per process: this is the fused path again unless the env was set at start)
    w1, b1, w2, b2 = (t.double() for t in (ff.squeeze.weight, ff.squeeze.bias, ff.unsqueeze.weight, ff.unsqueeze.bias))
    ref = (x.double() @ w1.t() + b1).relu() @ w2.t() + b2
    assert_close(out, ref.float(), 1e-3, 1e-4, "fused FFN vs fp64")
    assert_close(out, ref2, 1e-3, 1e-4, "fused FFN vs second call")


# ----------------------------------------------------------------------------------------------- conv front-end
@pytest.mark.parametrize("B,Fdim,T", [(2, 80, 200), (1, 80, 1000), (3, 33, 71), (1, 513, 311)])
def test_conv_frontend(B, Fdim, T):
    import asr_transformer_b200 as A
    torch.manual_seed(1)
    front = A.ConvFrontEnd(torch.nn.Conv2d(1, 64, 3, stride=2), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, stride=2),
                           torch.nn.ReLU())
    with torch.no_grad():
        for p in front.parameters():
            O.bf16_representable_(p)
    spec = O.structured_spectrum(B, T, Fdim, seed=2)
    sd = {"input_layer." + k: v for k, v in front.state_dict().items()}
    ref = O.frontend(sd, spec)
    out = front.to(DEV)(spec.to(DEV))
    assert out.shape == ref.shape
    assert_close(out, ref, 6e-2, 4e-3, "conv front-end")    # conv1 output and conv2 result are rounded to bf16


@pytest.mark.parametrize("B,Fdim,T", [(2, 80, 200), (3, 80, 1000), (3, 33, 71), (1, 513, 311), (2, 80, 7), (1, 11, 45)])
def test_conv_fused_equals_split(B, Fdim, T, monkeypatch):
    """The fused front-end (conv1 patch kept in shared memory) and the two-kernel path do the same arithmetic in the
    same order: bit-identical outputs, including ragged last tiles and inputs wider than one tile."""
    import asr_transformer_b200 as A
    torch.manual_seed(3)
    front = A.ConvFrontEnd(torch.nn.Conv2d(1, 64, 3, stride=2), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, stride=2),
                           torch.nn.ReLU()).to(DEV)
    spec = O.structured_spectrum(B, T, Fdim, seed=4).to(DEV)
    monkeypatch.setenv("ASR_B200_CONV", "split")
    ref = front(spec)
    monkeypatch.setenv("ASR_B200_CONV", "fused")
    out = front(spec)
    sync()
    assert out.shape == ref.shape and torch.equal(out, ref)


@pytest.mark.parametrize("B,T", [(2, 200), (3, 1000), (1, 15), (4, 301), (2, 7)])
def test_conv_tc_matches_fused(B, T, monkeypatch):
    """The tcgen05 front-end (conv_tc.cu: conv2 as an implicit GEMM over parity planes of the conv1 patch) against the
    mma.sync fused kernel: same operands (fp32 conv1, f16 hi | lo split into conv2), another fp32 summation order, so
    equal to fp32 rounding noise; both against the oracle at the front-end tolerance.  Ragged last tiles included."""
    import asr_transformer_b200 as A
    torch.manual_seed(3)
    front = A.ConvFrontEnd(torch.nn.Conv2d(1, 64, 3, stride=2), torch.nn.ReLU(), torch.nn.Conv2d(64, 64, 3, stride=2),
                           torch.nn.ReLU())
    with torch.no_grad():
        for p in front.parameters():
            O.bf16_representable_(p)
    spec = O.structured_spectrum(B, T, 80, seed=4)
    sd = {"input_layer." + k: v for k, v in front.state_dict().items()}
    ref = O.frontend(sd, spec)
    front = front.to(DEV)
    monkeypatch.setenv("ASR_B200_CONV", "fused")
    legacy = front(spec.to(DEV))
    monkeypatch.setenv("ASR_B200_CONV", "")
    out = front(spec.to(DEV))
    sync()
    assert out.shape == legacy.shape == ref.shape
    assert_close(out, legacy, 2e-3, 1e-4, "tcgen05 vs mma.sync front-end")
    assert_close(out, ref, 6e-2, 4e-3, "tcgen05 front-end vs oracle")


def test_embed_pe():
    V, D, B, Ls = 250, 256, 3, 17
    emb, pe = rnd(V, D, seed=40), rnd(64, D, seed=41)
    tok = torch.randint(0, V, (B, Ls), generator=torch.Generator().manual_seed(1)).to(torch.int32).to(DEV)
    out = torch.empty(B, Ls, D, device=DEV)
    L.check(lib().asr_embed_pe(L.ptr(tok), L.ptr(emb), L.ptr(pe), B, Ls, D, V, L.ptr(out), L.stream()))
    sync()
    assert torch.equal(out, emb[tok.long()] + pe[:Ls])


# ----------------------------------------------------------------------------------------------- decode-step ops
@pytest.mark.parametrize("B,N,K,ln,relu,res", [(64, 768, 256, 1, 0, 0), (5, 256, 256, 0, 0, 1), (70, 1024, 256, 1, 1, 0),
                                               (64, 256, 1024, 0, 0, 1), (3, 250, 256, 0, 0, 0), (33, 1536, 512, 1, 0, 0)])
def test_dec_linear_fp32_accurate(B, N, K, ln, relu, res):
    x = rnd(B, K, seed=50, scale=2.0) + 0.3
    npad = (N + 63) // 64 * 64
    w = torch.zeros(npad, K, dtype=torch.float16, device=DEV)
    w[:N] = rnd(N, K, seed=51, scale=K ** -0.5, dtype=torch.float16)
    bias = rnd(N, seed=52)
    g, b = (rnd(K, seed=53) * 0.2 + 1.0, rnd(K, seed=54) * 0.1) if ln else (None, None)
    resid = rnd(B, N, seed=55) if res else None
    out = torch.zeros(B, N, device=DEV)
    L.check(lib().asr_dec_linear(L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(w), L.ptr(bias), L.ptr(resid), B, N, K, relu,
                                 L.ptr(out), L.stream()))
    sync()
    a = F.layer_norm(x, (K,), g, b, 1e-5) if ln else x
    ref = a.double() @ w[:N].double().t() + bias.double()
    if relu:
        ref = ref.relu()
    if res:
        ref = ref + resid.double()
    assert_close(out, ref.float(), 2e-4, 2e-5, "dec_linear (bf16 hi+lo split must be fp32-accurate)")


@pytest.mark.parametrize("B,H,n", [(64, 4, 249), (3, 2, 1), (5, 8, 749), (2, 4, 17)])
def test_dec_attention(B, H, n):
    D = 64 * H
    q = rnd(B, D, seed=60)
    kv = rnd(B, n, 2 * D, seed=61, dtype=torch.float16)
    out = torch.zeros(B, D, device=DEV)
    scale = D ** -0.5
    L.check(lib().asr_dec_attention(L.ptr(q), L.ptr(kv), kv.data_ptr() + 2 * D, 2 * D, n * 2 * D, n, B, H, scale,
                                    L.ptr(out), L.stream()))
    sync()
    k = kv[:, :, :D].float().view(B, n, H, 64)
    v = kv[:, :, D:].float().view(B, n, H, 64)
    s = torch.einsum("bhd,bkhd->bhk", q.view(B, H, 64), k) * scale
    ref = torch.einsum("bhk,bkhd->bhd", torch.softmax(s, -1), v).reshape(B, D)
    assert_close(out, ref, 1e-4, 1e-5, "dec_attention")
