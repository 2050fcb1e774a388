"""fp32 CPU restatement of the reference Speech-Transformer hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function works on
a reference-keyed ``state_dict`` with plain torch CPU
tensor ops in fp32; citations are ``file:line`` relative to the reference tree
(``modules/Transformer/...``).

Two greedy decoders are provided on purpose:

* :func:`evaluate_reference_style` follows ``model.py:125-151`` literally: one
  utterance at a time, no KV cache, the whole prefix and the cross-attention
  K/V projections recomputed every step, exactly ``L`` steps, no stop at EOS.
  It is what ``bench.py --impl reference`` / ``cpu_baseline`` times, because
  that *is* the reference's CPU implementation of the path.
* :func:`greedy_kv_cached` is the batched, KV-cached restatement (token-exact in
  fp32) used as the fast checker at BASELINE sizes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------
# configuration registry and synthetic inputs: shared with bench.py / the tools,
# defined in the package (pure host-side data, no kernels)
# --------------------------------------------------------------------------
from asr_transformer_b200.workloads import (CONFIGS, Config, bf16_representable_, conv_len,  # noqa: E402,F401
                                            structured_spectrum, subsampled_len)


def teacher_text(cfg: Config, batch: int, seed: int = 2) -> Tuple[Tensor, Tensor]:
    """Random text (B,L) int64 and mask (B,L) (>=1 real token) with padded tails."""
    g = torch.Generator().manual_seed(seed)
    L = cfg.decoder_seq_len
    text = torch.randint(5, cfg.vocab_size, (batch, L), generator=g)
    lens = torch.randint(max(1, L // 2), L + 1, (batch,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.int64)
    text = torch.where(mask > 0, text, torch.full_like(text, cfg.pad_token_id))
    return text, mask


# --------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------
def positional_encoding(seq_len: int, emb_dim: int) -> Tensor:
    """layers.py:61-73: arg[p,j] = p / 10000^(j/D), j=0..D-1; first half sin, second half cos."""
    pos = torch.arange(0, seq_len).unsqueeze(1).float()
    arg = pos / (10000.0 ** (torch.arange(0, emb_dim).float() / emb_dim))
    pe = torch.zeros(seq_len, emb_dim)
    h = emb_dim // 2
    pe[:, :h] = torch.sin(arg[:, :h])
    pe[:, h:] = torch.cos(arg[:, h:])
    return pe.unsqueeze(0)


def layer_norm(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def linear(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def mha(sd: StateDict, prefix: str, x: Tensor, enc_x: Optional[Tensor] = None,
        attention_mask: Optional[Tensor] = None) -> Tensor:
    """layers.py:15-40.  Scale is emb_dim**-0.5 (d_model, layers.py:20); fully masked
    rows give zeros through nan_to_num (layers.py:25); heads concatenated in index order."""
    emb_dim = x.shape[-1]
    src = x if enc_x is None else enc_x
    heads = []
    h = 0
    while f"{prefix}._heads.{h}._q.weight" in sd:
        hp = f"{prefix}._heads.{h}"
        v = linear(sd, hp + "._v", src)
        k = linear(sd, hp + "._k", src)
        q = linear(sd, hp + "._q", x)
        s = q.bmm(k.transpose(1, 2)) * (emb_dim ** (-0.5))
        if attention_mask is not None:
            s = s.masked_fill(attention_mask.gt(0), float("-inf"))
        p = torch.nan_to_num(torch.softmax(s, dim=-1))
        heads.append(p.bmm(v))
        h += 1
    return linear(sd, prefix + "._out_linear", torch.cat(heads, dim=-1))


def feed_forward(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """layers.py:53-58 (dropout = identity in eval)."""
    return linear(sd, prefix + ".unsqueeze", torch.relu(linear(sd, prefix + ".squeeze", x)))


def frontend(sd: StateDict, spectrum: Tensor) -> Tensor:
    """model.py:168-171: two valid 3x3 stride-2 convs with ReLU. (B,1,F,T)->(B,64,F',T')."""
    y = torch.relu(F.conv2d(spectrum, sd["input_layer.0.weight"], sd["input_layer.0.bias"], stride=2))
    return torch.relu(F.conv2d(y, sd["input_layer.2.weight"], sd["input_layer.2.bias"], stride=2))


def encoder_layer(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """model.py:18-25 (pre-LN; _norm_in is never applied)."""
    x = mha(sd, prefix + "._attention", layer_norm(sd, prefix + "._norm1", x)) + x
    return feed_forward(sd, prefix + "._feedforward", layer_norm(sd, prefix + "._norm2", x)) + x


def encoder(sd: StateDict, y: Tensor, prefix: str = "encoder") -> Tensor:
    """model.py:41-52.  y: (B,64,F',T') -> (B,T',D)."""
    B, C, Fp, Tp = y.shape
    x = y.reshape(B, C * Fp, Tp).transpose(1, 2).contiguous()
    x = linear(sd, prefix + "._lin_in", x) + sd[prefix + "._pe.pe"][:, :Tp]
    i = 0
    while f"{prefix}._layers.{i}._norm1.weight" in sd:
        x = encoder_layer(sd, f"{prefix}._layers.{i}", x)
        i += 1
    return layer_norm(sd, prefix + "._norm_out", x)


def decoder_layer(sd: StateDict, prefix: str, x: Tensor, mask: Optional[Tensor], enc_x: Tensor) -> Tensor:
    """model.py:65-75: masked self-MHA, unmasked cross-MHA, FFN; each pre-LN + residual."""
    x = mha(sd, prefix + "._mask_attention", layer_norm(sd, prefix + "._norm1", x), None, mask) + x
    x = mha(sd, prefix + "._cross_attention", layer_norm(sd, prefix + "._norm2", x), enc_x) + x
    return feed_forward(sd, prefix + "._feedforward", layer_norm(sd, prefix + "._norm3", x)) + x


def _num_layers(sd: StateDict, prefix: str) -> int:
    i = 0
    while f"{prefix}._layers.{i}._norm1.weight" in sd:
        i += 1
    return i


def decoder_forward(sd: StateDict, text: Tensor, mask: Tensor, enc_x: Tensor, prefix: str = "decoder") -> Tensor:
    """model.py:104-123.  masked[b,i,j] = pad[b,j] or pad[b,i] or j>i; logits = classifier(LN(h))."""
    B, L = text.shape
    pad = mask.lt(1)
    m = pad.unsqueeze(1).expand(-1, L, -1)
    causal = torch.triu(torch.ones(L, L, dtype=torch.uint8), diagonal=1).bool().unsqueeze(0)
    m = m | m.transpose(1, 2) | causal
    x = F.embedding(text.long(), sd[prefix + "._embedding.weight"]) + sd[prefix + "._pe.pe"][:, :L]
    for i in range(_num_layers(sd, prefix)):
        x = decoder_layer(sd, f"{prefix}._layers.{i}", x, m, enc_x)
    x = layer_norm(sd, prefix + "._norm_layer", x)
    return F.linear(x, sd[prefix + "._classifier.weight"])


def transformer_forward(sd: StateDict, spectrum: Tensor, text: Tensor, mask: Tensor) -> Tensor:
    """model.py:194-198."""
    return decoder_forward(sd, text, mask, encoder(sd, frontend(sd, spectrum)))


def encode(sd: StateDict, spectrum: Tensor) -> Tensor:
    return encoder(sd, frontend(sd, spectrum))


# --------------------------------------------------------------------------
# greedy decoding
# --------------------------------------------------------------------------
def decoder_evaluate_reference_style(sd: StateDict, x: Tensor, enc_x: Tensor, seq_len: int, eos: int,
                                     prefix: str = "decoder") -> Tuple[Tensor, List[Tensor], Tensor]:
    """model.py:125-151 restated literally (per sample, no KV cache, no final LN, no break).

    Returns (decoder_input of the LAST sample (1,L+1) int64, probs list, all tokens (B,L+1) int64);
    the third item is extra (the reference builds and discards it, model.py:128,148).
    """
    B = x.shape[0]
    probs: List[Tensor] = []
    rows = []
    emb = sd[prefix + "._embedding.weight"]
    pe = sd[prefix + "._pe.pe"]
    nl = _num_layers(sd, prefix)
    dec_in = None
    for b in range(B):
        dec_in = x[b].unsqueeze(0)
        enc_b = enc_x[b].unsqueeze(0)
        for i in range(1, seq_len + 1):
            causal = torch.triu(torch.ones(i, i, dtype=torch.uint8), diagonal=1)
            h = F.embedding(dec_in.long(), emb) + pe[:, :i]
            for l in range(nl):
                h = decoder_layer(sd, f"{prefix}._layers.{l}", h, causal, enc_b)
            prob = F.linear(h, sd[prefix + "._classifier.weight"])      # model.py:142: NO _norm_layer
            nxt = prob.argmax(dim=-1)[:, -1].unsqueeze(1)
            dec_in = torch.cat([dec_in, nxt], dim=-1)                     # int32 + int64 -> int64
            if nxt.item() == eos or i == seq_len:
                probs.append(prob[:, :-1].squeeze())
        rows.append(dec_in)
    return dec_in, probs, torch.cat(rows, dim=0)


def evaluate_reference_style(sd: StateDict, spectrum: Tensor, cfg: Config):
    """model.py:201-206 with the BOS tensor of train.py:70."""
    enc_x = encode(sd, spectrum)
    bos = torch.full((spectrum.shape[0], 1), cfg.bos_token_id, dtype=torch.int32)
    return decoder_evaluate_reference_style(sd, bos, enc_x, cfg.decoder_seq_len, cfg.eos_token_id)


def _packed_heads(sd: StateDict, prefix: str, which: str) -> Tuple[Tensor, Tensor]:
    ws, bs = [], []
    h = 0
    while f"{prefix}._heads.{h}.{which}.weight" in sd:
        ws.append(sd[f"{prefix}._heads.{h}.{which}.weight"])
        bs.append(sd[f"{prefix}._heads.{h}.{which}.bias"])
        h += 1
    return torch.cat(ws, 0), torch.cat(bs, 0)


def greedy_kv_cached(sd: StateDict, enc_x: Tensor, cfg: Config, max_len: Optional[int] = None,
                     prefix: str = "decoder") -> Tuple[Tensor, Tensor]:
    """Batched KV-cached restatement of model.py:125-151 (exactly L steps, no final LN).

    Pre-LN + causal mask means row t of every layer depends only on rows <= t, so cached
    K/V equal the recomputed ones.  Returns (tokens (B,L+1) int64,
    step_logits (B,L,V) fp32) where step_logits[:,t] produced tokens[:,t+1].
    """
    B, Tp, D = enc_x.shape
    H = cfg.num_heads
    dh = D // H
    L = max_len or cfg.decoder_seq_len
    nl = _num_layers(sd, prefix)
    scale = D ** (-0.5)
    emb = sd[prefix + "._embedding.weight"]
    pe = sd[prefix + "._pe.pe"][0]
    Wc = sd[prefix + "._classifier.weight"]

    def heads(t):  # (B,S,D) -> (B,H,S,dh)
        return t.view(B, -1, H, dh).transpose(1, 2)

    packs = []
    for l in range(nl):
        lp = f"{prefix}._layers.{l}"
        sq, sk, sv = (_packed_heads(sd, lp + "._mask_attention", n) for n in ("_q", "_k", "_v"))
        cq, ck, cv = (_packed_heads(sd, lp + "._cross_attention", n) for n in ("_q", "_k", "_v"))
        ck_x = heads(F.linear(enc_x, *ck))
        cv_x = heads(F.linear(enc_x, *cv))
        packs.append((lp, sq, sk, sv, cq, ck_x, cv_x))

    tokens = torch.full((B, L + 1), cfg.bos_token_id, dtype=torch.int64)
    logits_all = torch.zeros(B, L, Wc.shape[0])
    kc = [torch.zeros(B, H, L, dh) for _ in range(nl)]
    vc = [torch.zeros(B, H, L, dh) for _ in range(nl)]
    for t in range(L):
        h = emb[tokens[:, t]] + pe[t]
        for l, (lp, sq, sk, sv, cq, ck_x, cv_x) in enumerate(packs):
            a = layer_norm(sd, lp + "._norm1", h)
            q = F.linear(a, *sq).view(B, H, 1, dh)
            kc[l][:, :, t] = F.linear(a, *sk).view(B, H, dh)
            vc[l][:, :, t] = F.linear(a, *sv).view(B, H, dh)
            s = (q @ kc[l][:, :, :t + 1].transpose(2, 3)) * scale
            o = torch.softmax(s, -1) @ vc[l][:, :, :t + 1]
            h = linear(sd, lp + "._mask_attention._out_linear", o.transpose(1, 2).reshape(B, D)) + h
            a = layer_norm(sd, lp + "._norm2", h)
            q = F.linear(a, *cq).view(B, H, 1, dh)
            s = (q @ ck_x.transpose(2, 3)) * scale
            o = torch.softmax(s, -1) @ cv_x
            h = linear(sd, lp + "._cross_attention._out_linear", o.transpose(1, 2).reshape(B, D)) + h
            h = feed_forward(sd, lp + "._feedforward", layer_norm(sd, lp + "._norm3", h)) + h
        lg = F.linear(h, Wc)
        logits_all[:, t] = lg
        tokens[:, t + 1] = lg.argmax(-1)
    return tokens, logits_all


def beam_search_kv_cached(sd: StateDict, enc_x: Tensor, cfg: Config, beam: int, max_len: Optional[int] = None,
                          prefix: str = "decoder") -> Tuple[Tensor, Tensor]:
    """Beam search on the step of :func:`greedy_kv_cached` (same logits: no final LayerNorm, model.py:142).

    The reference has NO beam search (README.md:30 lists it as a TODO), so this function DEFINES the semantics the CUDA
    path is held to ("parity unpinned" by the reference for this row):
      * every utterance keeps ``beam`` hypotheses; step 0 starts from one live hypothesis [BOS] with score 0, the others
        are dead (score -inf);
      * candidate score = hypothesis score + log_softmax(logits)[token], no length normalisation; a hypothesis that has
        emitted EOS is finished: its only candidate is itself (same score), extended with the pad token;
      * the ``beam`` best of the beam x V candidates survive, ties broken by the lower flat index (beam-major, then
        token) - the lowest-index rule of ``argmax`` (model.py:143) carried over; exactly L steps run.
    ``beam == 1`` is greedy search that pads after EOS.  Returns (tokens (B, beam, L+1) int64 best first, scores
    (B, beam) fp32)."""
    B, Tp, D = enc_x.shape
    K = int(beam)
    R = B * K
    H = cfg.num_heads
    dh = D // H
    L = max_len or cfg.decoder_seq_len
    nl = _num_layers(sd, prefix)
    scale = D ** (-0.5)
    emb = sd[prefix + "._embedding.weight"]
    pe = sd[prefix + "._pe.pe"][0]
    Wc = sd[prefix + "._classifier.weight"]
    V = Wc.shape[0]
    enc_r = enc_x.repeat_interleave(K, 0)

    def heads(t):
        return t.view(R, -1, H, dh).transpose(1, 2)

    packs = []
    for l in range(nl):
        lp = f"{prefix}._layers.{l}"
        sq, sk, sv = (_packed_heads(sd, lp + "._mask_attention", n) for n in ("_q", "_k", "_v"))
        cq, ck, cv = (_packed_heads(sd, lp + "._cross_attention", n) for n in ("_q", "_k", "_v"))
        packs.append((lp, sq, sk, sv, cq, heads(F.linear(enc_r, *ck)), heads(F.linear(enc_r, *cv))))

    tokens = torch.full((R, L + 1), cfg.bos_token_id, dtype=torch.int64)
    score = torch.full((B, K), float("-inf"))
    score[:, 0] = 0.0
    finished = torch.zeros(B, K, dtype=torch.bool)
    kc = [torch.zeros(R, H, L, dh) for _ in range(nl)]
    vc = [torch.zeros(R, H, L, dh) for _ in range(nl)]
    base = (torch.arange(B) * K)[:, None]
    for t in range(L):
        h = emb[tokens[:, t]] + pe[t]
        for l, (lp, sq, sk, sv, cq, ck_x, cv_x) in enumerate(packs):
            a = layer_norm(sd, lp + "._norm1", h)
            q = F.linear(a, *sq).view(R, H, 1, dh)
            kc[l][:, :, t] = F.linear(a, *sk).view(R, H, dh)
            vc[l][:, :, t] = F.linear(a, *sv).view(R, H, dh)
            s = (q @ kc[l][:, :, :t + 1].transpose(2, 3)) * scale
            o = torch.softmax(s, -1) @ vc[l][:, :, :t + 1]
            h = linear(sd, lp + "._mask_attention._out_linear", o.transpose(1, 2).reshape(R, D)) + h
            a = layer_norm(sd, lp + "._norm2", h)
            q = F.linear(a, *cq).view(R, H, 1, dh)
            s = (q @ ck_x.transpose(2, 3)) * scale
            o = torch.softmax(s, -1) @ cv_x
            h = linear(sd, lp + "._cross_attention._out_linear", o.transpose(1, 2).reshape(R, D)) + h
            h = feed_forward(sd, lp + "._feedforward", layer_norm(sd, lp + "._norm3", h)) + h
        logp = torch.log_softmax(F.linear(h, Wc), -1).view(B, K, V)
        cand = score[:, :, None] + logp
        keep = torch.full((B, K, V), float("-inf"))
        keep[:, :, cfg.pad_token_id] = score
        cand = torch.where(finished[:, :, None], keep, cand).view(B, K * V)
        val, idx = torch.sort(cand, dim=1, descending=True, stable=True)       # stable: lower flat index wins ties
        val, idx = val[:, :K], idx[:, :K]
        parent, tok = idx // V, idx % V
        rows = (base + parent).reshape(-1)
        tokens = tokens[rows]
        tokens[:, t + 1] = tok.reshape(-1)
        for l in range(nl):
            kc[l] = kc[l][rows]
            vc[l] = vc[l][rows]
        finished = torch.gather(finished, 1, parent) | (tok == cfg.eos_token_id)
        score = val
    return tokens.view(B, K, L + 1), score


def greedy(sd: StateDict, spectrum: Tensor, cfg: Config) -> Tuple[Tensor, Tensor]:
    return greedy_kv_cached(sd, encode(sd, spectrum), cfg)


# --------------------------------------------------------------------------
# near-tie tracer
# --------------------------------------------------------------------------
def compare_tokens(ref_tokens: Tensor, ref_step_logits: Tensor, got_tokens: Tensor, tau: float = 2e-2) -> dict:
    """Utterance-level token agreement; each divergence is classified by the fp32
    reference top1-top2 margin at the first differing step (near-tie iff margin < tau)."""
    ref_tokens = ref_tokens.long().cpu()
    got_tokens = got_tokens.long().cpu()
    B = ref_tokens.shape[0]
    identical, near_tie, hard = 0, [], []
    for b in range(B):
        diff = (ref_tokens[b] != got_tokens[b]).nonzero()
        if diff.numel() == 0:
            identical += 1
            continue
        pos = int(diff[0])            # token index; produced by step pos-1
        top2 = ref_step_logits[b, pos - 1].topk(2).values
        margin = float(top2[0] - top2[1])
        (near_tie if margin < tau else hard).append((b, pos, margin))
    return {"utterances": B, "identical": identical, "near_tie": near_tie, "hard": hard,
            "distinct_rows": len({tuple(r.tolist()) for r in ref_tokens})}
