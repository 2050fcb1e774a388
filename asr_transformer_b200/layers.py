"""Drop-in replacements for reference ``modules/Transformer/layers.py``.

Same class names, constructor signatures and ``state_dict`` keys; ``forward`` runs hand-written sm_100a kernels
through the C-ABI (``include/asr_b200.h``).  Inference only: dropout is the identity (the reference's eval path,
train.py:60), and there is no autograd through the kernels.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import lib as _l
from .engine import pack_ffn, pack_mha, weights_version, _f16, _f32


def _require_eval(m: nn.Module) -> None:
    if m.training:
        raise RuntimeError(
            f"{type(m).__name__}: asr_transformer_b200 implements the inference path only (no dropout, no backward); "
            "call model.eval() first, as the reference's eval_epoch does (train.py:60).")


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm call sites of the reference (model.py:14,16,33,59,61,63,101): same parameters, CUDA kernel."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        D = self.normalized_shape[0]
        x = x.to(torch.float32).contiguous()
        y = torch.empty_like(x)
        with _l.on(x):   # (a half-precision module would otherwise be read as fp32: ptr() checks the dtype)
            _l.check(_l.load().asr_layernorm(_l.ptr(x), _l.ptr(self.weight.detach(), torch.float32),
                                             _l.ptr(self.bias.detach(), torch.float32), x.numel() // D, D, _l.ptr(y), None,
                                             _l.stream()), "asr_layernorm")
        return y


class MHAHead(nn.Module):
    """reference layers.py:6-28. One attention head; scale is emb_dim ** -0.5 (layers.py:20)."""

    def __init__(self, emb_dim, head_dim, dropout):
        super().__init__()
        self._emb_dim = emb_dim
        self._v = nn.Linear(emb_dim, head_dim)
        self._q = nn.Linear(emb_dim, head_dim)
        self._k = nn.Linear(emb_dim, head_dim)
        self._dropout = nn.Dropout(dropout)

    def forward(self, x, enc_x=None, attention_mask=None):
        _require_eval(self)
        if self._q.out_features != 64 or self._emb_dim % 64 != 0:
            raise RuntimeError("asr_b200: only head_dim 64 and emb_dim % 64 == 0 are implemented")
        L = _l.load()
        src = x if enc_x is None else enc_x
        B, Sq, D = x.shape
        Sk = src.shape[1]
        dev = x.device
        w = _f16(torch.cat([self._q.weight, self._k.weight, self._v.weight], 0))       # [192, D]
        b = _f32(torch.cat([self._q.bias, self._k.bias, self._v.bias], 0))
        xb = torch.empty(B * Sq, D, dtype=torch.float16, device=dev)
        xf = x.to(torch.float32).contiguous()
        _l.check(L.asr_f32_to_f16(_l.ptr(xf), _l.ptr(xb), xb.numel(), _l.stream()))
        q = torch.empty(B * Sq, 64, dtype=torch.float16, device=dev)
        kv = torch.empty(B * Sk, 128, dtype=torch.float16, device=dev)
        _l.check(L.asr_gemm_f16(_l.ptr(xb), _l.ptr(w), _l.ptr(b), None, None, 1, B * Sq, 64, D, 0, None, _l.ptr(q), 0,
                                 _l.stream()))
        if enc_x is None:
            sb = xb
        else:
            sf = src.to(torch.float32).contiguous()
            sb = torch.empty(B * Sk, D, dtype=torch.float16, device=dev)
            _l.check(L.asr_f32_to_f16(_l.ptr(sf), _l.ptr(sb), sb.numel(), _l.stream()))
        wkv, bkv = w[64:].contiguous(), b[64:].contiguous()
        _l.check(L.asr_gemm_f16(_l.ptr(sb), _l.ptr(wkv), _l.ptr(bkv), None, None, 1, B * Sk, 128, D, 0, None,
                                 _l.ptr(kv), 0, _l.stream()))
        dense, mask_b = _dense_mask(attention_mask, B, Sq, Sk, dev)
        out = torch.empty(B, Sq, 64, dtype=torch.float16, device=dev)
        _l.check(L.asr_attention(_l.ptr(q), 64, Sq * 64, _l.ptr(kv), 128, Sk * 128, kv.data_ptr() + 128, 128, Sk * 128,
                                 _l.ptr(out), 64, Sq * 64, B, 1, Sq, Sk, float(self._emb_dim) ** -0.5, 0, None, None,
                                 None, _l.ptr(dense), mask_b, 0, _l.stream()), "asr_attention")
        return out.float()


def _dense_mask(attention_mask, B, Sq, Sk, dev):
    """reference layers.py:22-23: masked where attention_mask > 0; (B,Sq,Sk) or broadcastable (Sq,Sk)."""
    if attention_mask is None:
        return None, 1
    m = attention_mask.to(dev).gt(0)
    if m.dim() == 2:
        m = m.unsqueeze(0)
    if m.shape[-2:] != (Sq, Sk) or m.shape[0] not in (1, B):
        m = m.expand(B, Sq, Sk)
    return m.to(torch.uint8).contiguous(), m.shape[0]


class MHA(nn.Module):
    """reference layers.py:31-40: heads concatenated in index order, then _out_linear."""

    def __init__(self, num_heads, emb_dim, dropout):
        super().__init__()
        self._dropout = nn.Dropout(dropout)
        self._heads = nn.ModuleList([MHAHead(emb_dim, emb_dim // num_heads, dropout) for _ in range(num_heads)])
        self._out_linear = nn.Linear(emb_dim, emb_dim)
        self._packed = None
        self._packed_version = None

    def _pack(self):
        ver = weights_version(self)
        if ver != self._packed_version:
            self._packed = pack_mha(self)
            self._packed_version = ver
        return self._packed

    def forward(self, x, enc_x=None, attention_mask=None):
        _require_eval(self)
        L = _l.load()
        p = self._pack()
        B, Sq, D = x.shape
        H = len(self._heads)
        Sk = Sq if enc_x is None else enc_x.shape[1]
        dev = x.device
        xf = x.to(torch.float32).contiguous()
        sf = None if enc_x is None else enc_x.to(torch.float32).contiguous()
        dense, mask_b = _dense_mask(attention_mask, B, Sq, Sk, dev)
        out = torch.empty(B, Sq, D, dtype=torch.float32, device=dev)
        ws = _l.workspace(L.asr_mha_workspace_bytes(B, Sq, Sk, D), dev, "op")
        wts = _l.AsrMhaWeights(p["w_qkv"].data_ptr(), p["b_qkv"].data_ptr(), p["w_out"].data_ptr(),
                               p["b_out"].data_ptr())
        _l.check(L.asr_mha(_l.ptr(xf), _l.ptr(sf), C.byref(wts), B, Sq, Sk, D, H, 0, None, None, _l.ptr(dense), mask_b,
                           _l.ptr(ws), ws.numel(), _l.ptr(out), _l.stream()), "asr_mha")
        return out


class FeedForward(nn.Module):
    """reference layers.py:43-58."""

    def __init__(self, emb_dim, ff_dim, dropout):
        super().__init__()
        self.emb_dim = emb_dim
        self.ff_dim = ff_dim
        self.squeeze = nn.Linear(self.emb_dim, self.ff_dim)
        self.ReLU = nn.ReLU()
        self.dropout = nn.Dropout(dropout)
        self.unsqueeze = nn.Linear(self.ff_dim, self.emb_dim)
        self._packed = None
        self._packed_version = None

    def forward(self, x):
        _require_eval(self)
        L = _l.load()
        ver = weights_version(self)
        if ver != self._packed_version:
            self._packed, self._packed_version = pack_ffn(self), ver
        p = self._packed
        xf = x.to(torch.float32).contiguous()
        rows = xf.numel() // self.emb_dim
        out = torch.empty_like(xf)
        ws = _l.workspace(L.asr_ffn_workspace_bytes(rows, self.emb_dim, self.ff_dim), xf.device, "op")
        wts = _l.AsrFfnWeights(p["w1"].data_ptr(), p["b1"].data_ptr(), p["w2"].data_ptr(), p["b2"].data_ptr())
        _l.check(L.asr_ffn(_l.ptr(xf), C.byref(wts), rows, self.emb_dim, self.ff_dim, _l.ptr(ws), ws.numel(),
                           _l.ptr(out), _l.stream()), "asr_ffn")
        return out


class TrainablePositionalEncoding(nn.Module):
    """reference layers.py:61-73: a fixed buffer (kept in state_dict); arg[p,j] = p / 10000^(j/D) for j = 0..D-1,
    first half sin, second half cos."""

    def __init__(self, seq_len, emb_dim):
        super().__init__()
        half = emb_dim // 2
        pos = torch.arange(0, seq_len).unsqueeze(1).float()
        arg = pos / (10000. ** (torch.arange(0, emb_dim).float() / emb_dim))      # exponent j/D, j = 0..D-1
        pe = torch.cat([torch.sin(arg[:, :half]), torch.cos(arg[:, half:])], dim=1)   # half split, not interleaved
        self.register_buffer('pe', pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, :x.size(1)]
