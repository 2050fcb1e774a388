"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the weight
packer is exact, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import asr_transformer_b200 as A
from asr_transformer_b200 import engine as E
from asr_transformer_b200 import lib as L
from oracle import speech_transformer as O
from tests.util import build_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/asr_b200.h but not exported"
    assert sorted(L._SIGNATURES.keys()) == names, "ctypes signature table out of sync with the header"
    assert lib.asr_version() == 1


def test_create_rejects_unsupported_config_without_gpu():
    lib = L.load()
    cfg = L.AsrConfig(250, 80, 96, 16, 49, 1, 1, 2, 256, 4, 2, 1)     # head_dim 48
    h = ctypes.c_void_p()
    rc = lib.asr_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and b"head_dim" in lib.asr_last_error()
    cfg = L.AsrConfig(250, 80, 128, 16, 49, 1, 1, 2, 256, 4, 2, 1)
    assert lib.asr_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    n = ctypes.c_size_t()
    assert lib.asr_workspace_bytes(h, 3, 200, 16, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.asr_encode(h, None, 1, 200, None, None, 0, None, None) == -1     # weights not loaded
    lib.asr_destroy(h)


def test_state_dict_keys_and_dead_parameters():
    m = build_model(O.CONFIGS["T0"])
    keys = list(m.state_dict().keys())
    assert "input_encoding.weight" in keys and "encoder._layers.0._norm_in.weight" in keys      # dead params (Q9)
    assert "encoder._pe.pe" in keys and "decoder._pe.pe" in keys
    assert "decoder._classifier.weight" in keys and "decoder._classifier.bias" not in keys
    assert "encoder._layers.1._attention._heads.1._k.bias" in keys
    assert (m.decoder._embedding.weight[4] == 0).all()                                         # padding_idx row
    m2 = build_model(O.CONFIGS["T0"])
    m2.load_state_dict(m.state_dict(), strict=True)


def test_packers_are_exact():
    m = build_model(O.CONFIGS["T0"])
    mha = m.encoder._layers[0]._attention
    p = E.pack_mha(mha)
    D = 128
    assert p["w_qkv"].shape == (3 * D, D) and p["w_qkv"].dtype == torch.float16
    # bf16-representable weights are exact in fp16 down to 2^-14; smaller magnitudes round by at most 2^-25
    for got, ref in ((p["w_qkv"][64:128].float(), mha._heads[1]._q.weight), (p["w_qkv"][D:D + 64].float(), mha._heads[0]._k.weight)):
        big = ref.abs() >= 2.0 ** -14
        assert torch.equal(got[big], ref[big]) and (got - ref).abs().max().item() <= 2.0 ** -25
    assert torch.equal(p["b_qkv"][2 * D + 64:], mha._heads[1]._v.bias)
    w = m.encoder._lin_in.weight
    pw = E.pack_lin_in(w).float()
    Fp = 19
    c, f = 37, 11
    assert (pw[:, f * 64 + c] - w[:, c * Fp + f]).abs().max().item() <= 2.0 ** -25
    w1 = m.input_layer[0].weight
    assert torch.equal(E.pack_conv1(w1)[1 * 3 + 2, 17], w1[17, 0, 1, 2])
    w2 = m.input_layer[2].weight
    frag = E.pack_conv2_fragments(w2).float()
    tap, half, sub, nt, g, cc, j = 5, 1, 0, 3, 6, 2, 3
    assert abs(float(frag[(tap * 2 + half) * 2 + sub, nt, g * 4 + cc, j]) -
               float(w2[nt * 8 + g, half * 32 + cc * 8 + sub * 4 + j, tap // 3, tap % 3])) <= 2.0 ** -25
    cw = E.pack_classifier(m.decoder._classifier.weight)
    assert cw.shape == (256, D) and (cw[250:] == 0).all()


def test_no_cpu_fallback():
    m = build_model(O.CONFIGS["T0"])
    spec = O.structured_spectrum(1, 200)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.greedy_decode(spec)
    with pytest.raises(RuntimeError, match="no CPU path"):
        A.LayerNorm(128).eval()(torch.zeros(2, 128))
    m.train()
    with pytest.raises(RuntimeError, match="inference path only"):
        m(spec, torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 4))


def test_positional_encoding_closed_form():
    pe = A.TrainablePositionalEncoding(50, 128).pe
    assert torch.equal(pe, O.positional_encoding(50, 128))
    assert pe.shape == (1, 50, 128)
