"""Weight packing and the per-model engine that owns the C-ABI handle.

Packing happens once per weight version:
  * per-head ``_q/_k/_v`` Linear(D, 64) -> one fp16 ``[3D, D]`` matrix (rows: q heads | k heads | v heads),
  * ``input_layer.0.weight`` -> fp32 ``[9 taps, 64]``; ``input_layer.2.weight`` -> fp16 mma-fragment order,
  * ``encoder._lin_in.weight`` columns permuted from the reference's ``c*F'+f`` (model.py:43-45) to ``f*64+c``,
    which is the order the conv kernel writes, so no transpose/contiguous pass exists at run time,
  * ``decoder._classifier.weight`` zero-padded to a multiple of 64 rows.
GEMM weights are stored as fp16 (lossless when the checkpoint is bf16-representable, as in the parity tests: fp16 has
three more mantissa bits; only magnitudes below 2^-14 round, by at most 2^-25).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import lib as _l


def conv_len(n: int) -> int:
    return (n - 3) // 2 + 1


def _f16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float16).contiguous()


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def pack_mha(mha: nn.Module) -> dict:
    """reference layers.py:10-12,35-36 -> packed [3D, D] / [3D] (+ out projection)."""
    heads = list(mha._heads)
    wq = torch.cat([h._q.weight for h in heads], 0)
    wk = torch.cat([h._k.weight for h in heads], 0)
    wv = torch.cat([h._v.weight for h in heads], 0)
    bq = torch.cat([h._q.bias for h in heads], 0)
    bk = torch.cat([h._k.bias for h in heads], 0)
    bv = torch.cat([h._v.bias for h in heads], 0)
    return {"w_qkv": _f16(torch.cat([wq, wk, wv], 0)), "b_qkv": _f32(torch.cat([bq, bk, bv], 0)),
            "w_out": _f16(mha._out_linear.weight), "b_out": _f32(mha._out_linear.bias)}


def pack_ffn(ff: nn.Module) -> dict:
    return {"w1": _f16(ff.squeeze.weight), "b1": _f32(ff.squeeze.bias),
            "w2": _f16(ff.unsqueeze.weight), "b2": _f32(ff.unsqueeze.bias)}


def pack_norm(ln: nn.Module) -> dict:
    return {"gamma": _f32(ln.weight), "beta": _f32(ln.bias)}


def pack_conv1(w: torch.Tensor) -> torch.Tensor:
    """(64,1,3,3) -> fp32 [tap = kh*3+kw][64]."""
    return _f32(w.detach().reshape(64, 9).t())


def pack_conv2_fragments(w: torch.Tensor) -> torch.Tensor:
    """(co=64, ci=64, kh, kw) -> fp16 [36 k-steps][8 n-tiles][32 lanes][4] in mma.m16n8k16 B-fragment order.

    k-step ks = (tap*2 + half)*2 + sub covers channels half*32 + c*8 + sub*4 + {0..3} for lane (g, c) = divmod(lane, 4)
    and output channel co = nt*8 + g (see conv2_kernel in csrc/simple_ops.cu)."""
    taps = w.detach().permute(2, 3, 0, 1).reshape(9, 64, 64)          # [tap, co, ci]
    t = taps.reshape(9, 8, 8, 2, 4, 2, 4)                              # tap, nt, g, half, c, sub, j
    t = t.permute(0, 3, 5, 1, 2, 4, 6)                                 # tap, half, sub, nt, g, c, j
    return _f16(t.reshape(36, 8, 32, 4))


def pack_lin_in(w: torch.Tensor) -> torch.Tensor:
    """(D, 64*F') columns c*F'+f -> f*64+c."""
    D, K = w.shape
    Fp = K // 64
    return _f16(w.detach().reshape(D, 64, Fp).permute(0, 2, 1).reshape(D, K))


def pack_classifier(w: torch.Tensor) -> torch.Tensor:
    V, D = w.shape
    vpad = (V + 63) // 64 * 64
    out = torch.zeros(vpad, D, dtype=torch.float16, device=w.device)
    out[:V] = w.detach().to(torch.float16)
    return out


def dec_image_layout(D: int, H: int, FF: int, V: int, nd: int) -> Optional[dict]:
    """Byte layout of the cluster decoder's packed image; mirrors cluster_layout() in csrc/decode_cluster.cu
    (documented in include/asr_b200.h).  None when the configuration is not supported by that kernel."""
    if H < 2 or H > 8 or (H & (H - 1)) or D != 64 * H or FF % (32 * H) or V < 1:
        return None
    FFS = FF // H
    VS = ((V + H - 1) // H + 15) // 16 * 16
    if D > 512 or FFS > 512 or VS > 512:
        return None
    lay = {"CS": H, "FFS": FFS, "VS": VS, "small_floats": 256 + FFS + 11 * D}
    lay["small_bytes"] = (lay["small_floats"] * 4 + 127) // 128 * 128
    if lay["small_bytes"] > 32768:
        return None
    off = 0
    for name, nbytes in (("small", lay["small_bytes"]), ("qkv", 192 * D * 2), ("wo", D * 64 * 2), ("wqc", 64 * D * 2),
                         ("woc", D * 64 * 2), ("w1", FFS * D * 2), ("w2", D * FFS * 2)):
        lay["off_" + name] = off
        off += nbytes
    lay["layer_bytes"] = off
    lay["off_cls"] = nd * off
    lay["rank_bytes"] = lay["off_cls"] + VS * D * 2
    lay["total_bytes"] = lay["rank_bytes"] * H
    return lay


def pack_mma_a(w: torch.Tensor) -> torch.Tensor:
    """fp16 [R, K] (R % 16 == 0, K % 32 == 0) -> flat fp16 in the cluster decoder's fragment-major order
    [k-block kb (32 cols)][m-tile mt (16 rows)][k-tile s (2)][g (8)][tg (4)][8], the 8 elements being the
    mma.m16n8k16 A fragment {a0, a1, a2, a3} of lane (g, tg): with r = 16 mt + 2 g and c = 32 kb + 8 tg + 4 s they are
    w[r, c:c+2], w[r+1, c:c+2], w[r, c+2:c+4], w[r+1, c+2:c+4] - one LDS.128 per lane feeds one MMA directly.  MMA rows
    g / g + 8 carry matrix rows 2g / 2g + 1 of the tile, so a lane's result fragment holds two ADJACENT output rows."""
    R, K = w.shape
    assert R % 16 == 0 and K % 32 == 0, (R, K)
    t = w.reshape(R // 16, 8, 2, K // 32, 4, 2, 2, 2)    # mt, g, p, kb, tg, s, pair, e
    return t.permute(3, 0, 5, 1, 4, 6, 2, 7).contiguous().reshape(-1)


def pack_dec_image(decoder: nn.Module) -> Optional[torch.Tensor]:
    """Packed weight image of the cluster decoder (uint8): for every head r (= CTA rank) the slices that CTA consumes,
    in consumption order, per layer: small fp32 block [b_qkv_r(192) | b_qc_r(64) | b1_r(FFS) | b_o(D) | b_oc(D) |
    b2(D) | ln1 g,b | ln2 g,b | ln3 g,b | ln1 g,b of the next layer] (padded to 128 B), then Wqkv rows of head r (q|k|v), Wo[:, head r columns],
    cross Wq rows of head r, cross Wo[:, head r columns], W1 rows r*FFS.., W2[:, r*FFS..] - each in pack_mma_a order;
    after the layers the classifier rows r*VS.. (zero padded)."""
    layers = list(decoder._layers)
    if not layers:
        return None
    D = decoder._embedding.embedding_dim
    V = decoder._embedding.num_embeddings
    H = len(layers[0]._mask_attention._heads)
    FF = layers[0]._feedforward.ff_dim
    lay = dec_image_layout(D, H, FF, V, len(layers))
    if lay is None:
        return None
    FFS, VS = lay["FFS"], lay["VS"]
    dev = decoder._embedding.weight.device
    cls = torch.zeros(H * VS, D, dtype=torch.float16, device=dev)
    cls[:V] = decoder._classifier.weight.detach().to(torch.float16)

    def u8(t: torch.Tensor) -> torch.Tensor:
        return t.contiguous().view(torch.uint8).reshape(-1)

    ranks = []
    for r in range(H):
        parts = []
        hs = slice(r * 64, (r + 1) * 64)
        for layer in layers:
            sa, ca, ff = pack_mha(layer._mask_attention), pack_mha(layer._cross_attention), pack_ffn(layer._feedforward)
            rows_qkv = torch.cat([torch.arange(r * 64, (r + 1) * 64) + k * D for k in range(3)]).to(dev)
            small = [sa["b_qkv"][rows_qkv], ca["b_qkv"][hs], ff["b1"][r * FFS:(r + 1) * FFS], sa["b_out"], ca["b_out"],
                     ff["b2"]]
            for ln in (layer._norm1, layer._norm2, layer._norm3):
                small += [_f32(ln.weight), _f32(ln.bias)]
            li = layers.index(layer)          # ... followed by norm1 of the NEXT layer (fused into this layer's last
            nxt = layers[li + 1]._norm1 if li + 1 < len(layers) else None          # all-reduce); zeros after the last
            small += [_f32(nxt.weight), _f32(nxt.bias)] if nxt is not None else [torch.zeros(2 * D, device=dev)]
            small = torch.cat([x.reshape(-1) for x in small])
            assert small.numel() == lay["small_floats"]
            small = torch.cat([small, small.new_zeros(lay["small_bytes"] // 4 - small.numel())])
            blk = [u8(small),
                   u8(pack_mma_a(sa["w_qkv"][rows_qkv])),
                   u8(pack_mma_a(sa["w_out"][:, hs])),
                   u8(pack_mma_a(ca["w_qkv"][hs])),
                   u8(pack_mma_a(ca["w_out"][:, hs])),
                   u8(pack_mma_a(ff["w1"][r * FFS:(r + 1) * FFS])),
                   u8(pack_mma_a(ff["w2"][:, r * FFS:(r + 1) * FFS]))]
            blk = torch.cat(blk)
            assert blk.numel() == lay["layer_bytes"]
            parts.append(blk)
        parts.append(u8(pack_mma_a(cls[r * VS:(r + 1) * VS])))
        rk = torch.cat(parts)
        assert rk.numel() == lay["rank_bytes"]
        ranks.append(rk)
    return torch.cat(ranks).contiguous()


def _mha_struct(p: dict) -> _l.AsrMhaWeights:
    return _l.AsrMhaWeights(p["w_qkv"].data_ptr(), p["b_qkv"].data_ptr(), p["w_out"].data_ptr(), p["b_out"].data_ptr())


def _ffn_struct(p: dict) -> _l.AsrFfnWeights:
    return _l.AsrFfnWeights(p["w1"].data_ptr(), p["b1"].data_ptr(), p["w2"].data_ptr(), p["b2"].data_ptr())


def _norm_struct(p: dict) -> _l.AsrNormWeights:
    return _l.AsrNormWeights(p["gamma"].data_ptr(), p["beta"].data_ptr())


def weights_version(module: nn.Module) -> Tuple:
    """Cheap fingerprint: any in-place update, .to(), load_state_dict or re-assignment changes it."""
    ps = list(module.parameters()) + list(module.buffers())
    return (len(ps), sum(p._version for p in ps), ps[0].data_ptr() if ps else 0, str(ps[0].device) if ps else "")


class Engine:
    """Owns an AsrHandle plus the packed device weights for (input_layer?, encoder?, decoder?)."""

    def __init__(self):
        self.handle = C.c_void_p()
        self.keep: List = []
        self.version = None
        self.cfg: Optional[_l.AsrConfig] = None
        self.device: Optional[torch.device] = None

    # the handle is process-local: copies / pickles start empty and re-pack on first use
    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    def __deepcopy__(self, memo):
        return Engine()

    def __del__(self):
        try:
            if self.handle:
                _l.load().asr_destroy(self.handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ build
    def sync(self, owner: nn.Module, input_layer=None, encoder=None, decoder=None, bos: int = 1) -> "Engine":
        ver = weights_version(owner)
        if ver == self.version:
            return self
        dev = next(owner.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("asr_b200 has no CPU path: move the model to a CUDA device (model.cuda())")
        L = _l.load()
        with _l.on(dev):
            return self._sync_on(owner, input_layer, encoder, decoder, bos, ver, dev, L)

    def _sync_on(self, owner, input_layer, encoder, decoder, bos, ver, dev, L) -> "Engine":
        keep: List = []
        w = _l.AsrWeights()
        cfg = _l.AsrConfig()
        D = H = FF = None
        if encoder is not None:
            D = encoder._lin_in.out_features
            Fp = encoder._lin_in.in_features // 64
            cfg.input_dim = 4 * Fp + 3 if input_layer is None else getattr(owner, "_input_dim", 4 * Fp + 3)
            cfg.encoder_seq_len = encoder._pe.pe.shape[1]
            cfg.encoder_num_layers = len(encoder._layers)
            lin_w, lin_b = pack_lin_in(encoder._lin_in.weight), _f32(encoder._lin_in.bias)
            pe = _f32(encoder._pe.pe[0])
            no = pack_norm(encoder._norm_out)
            keep += [lin_w, lin_b, pe, no]
            w.lin_in_w, w.lin_in_b, w.enc_pe = lin_w.data_ptr(), lin_b.data_ptr(), pe.data_ptr()
            w.enc_norm_out = _norm_struct(no)
            arr = (_l.AsrEncoderLayerWeights * max(1, len(encoder._layers)))()
            for i, layer in enumerate(encoder._layers):
                n1, a, n2, f = pack_norm(layer._norm1), pack_mha(layer._attention), pack_norm(layer._norm2), \
                    pack_ffn(layer._feedforward)
                keep += [n1, a, n2, f]
                arr[i] = _l.AsrEncoderLayerWeights(_norm_struct(n1), _mha_struct(a), _norm_struct(n2), _ffn_struct(f))
                H = len(layer._attention._heads)
                FF = layer._feedforward.ff_dim
            keep.append(arr)
            w.enc_layers = arr
        else:
            cfg.input_dim, cfg.encoder_seq_len, cfg.encoder_num_layers = 83, 1, 0
        if input_layer is not None:
            c1w, c1b = pack_conv1(input_layer[0].weight), _f32(input_layer[0].bias)
            c2w, c2b = pack_conv2_fragments(input_layer[2].weight), _f32(input_layer[2].bias)
            keep += [c1w, c1b, c2w, c2b]
            w.conv1_w, w.conv1_b, w.conv2_wfrag, w.conv2_b = c1w.data_ptr(), c1b.data_ptr(), c2w.data_ptr(), c2b.data_ptr()
        if decoder is not None:
            D = decoder._embedding.embedding_dim
            cfg.vocab_size = decoder._embedding.num_embeddings
            cfg.decoder_seq_len = decoder._pe.pe.shape[1]
            cfg.decoder_num_layers = len(decoder._layers)
            cfg.eos_token_id = int(decoder._eos_token_id)
            pad = decoder._embedding.padding_idx
            cfg.pad_token_id = int(pad) if pad is not None else 0
            emb, pe = _f32(decoder._embedding.weight), _f32(decoder._pe.pe[0])
            nl, cw = pack_norm(decoder._norm_layer), pack_classifier(decoder._classifier.weight)
            keep += [emb, pe, nl, cw]
            w.embedding, w.dec_pe, w.classifier_w = emb.data_ptr(), pe.data_ptr(), cw.data_ptr()
            w.dec_norm = _norm_struct(nl)
            arr = (_l.AsrDecoderLayerWeights * max(1, len(decoder._layers)))()
            for i, layer in enumerate(decoder._layers):
                n1, sa, n2 = pack_norm(layer._norm1), pack_mha(layer._mask_attention), pack_norm(layer._norm2)
                ca, n3, f = pack_mha(layer._cross_attention), pack_norm(layer._norm3), pack_ffn(layer._feedforward)
                keep += [n1, sa, n2, ca, n3, f]
                arr[i] = _l.AsrDecoderLayerWeights(_norm_struct(n1), _mha_struct(sa), _norm_struct(n2),
                                                   _mha_struct(ca), _norm_struct(n3), _ffn_struct(f))
                H = len(layer._mask_attention._heads)
                FF = layer._feedforward.ff_dim
            keep.append(arr)
            w.dec_layers = arr
            if len(decoder._layers):
                image = pack_dec_image(decoder)
                if image is not None:
                    keep.append(image)
                    w.dec_image, w.dec_image_bytes = image.data_ptr(), image.numel()
        else:
            cfg.vocab_size, cfg.decoder_seq_len, cfg.decoder_num_layers = 1, 1, 0
        if H is None:   # zero layers everywhere: derive from D
            H, FF = D // 64, 64
        cfg.embedding_dim, cfg.num_heads, cfg.ff_dim, cfg.bos_token_id = D, H, FF, bos
        if self.handle:
            L.asr_destroy(self.handle)
            self.handle = C.c_void_p()
        _l.check(L.asr_create(C.byref(cfg), C.byref(self.handle)), "asr_create")
        _l.check(L.asr_load_weights(self.handle, C.byref(w)), "asr_load_weights")
        self.keep, self.version, self.cfg, self.device = keep, ver, cfg, dev
        return self

    # ------------------------------------------------------------------ calls
    # every C call runs with the engine's device current (lib.on): the right stream, the right per-device kernel setup
    def encode(self, *a, **k):
        with _l.on(self.device):
            return self._encode_impl(*a, **k)

    def encoder_forward(self, *a, **k):
        with _l.on(self.device):
            return self._encoder_forward_impl(*a, **k)

    def decoder_forward(self, *a, **k):
        with _l.on(self.device):
            return self._decoder_forward_impl(*a, **k)

    def decode_greedy(self, *a, **k):
        with _l.on(self.device):
            return self._decode_greedy_impl(*a, **k)

    def decode_beam(self, *a, **k):
        with _l.on(self.device):
            return self._decode_beam_impl(*a, **k)

    def _ws(self, B: int, T: int, Ldec: int, tag: str = "model") -> torch.Tensor:
        """Scratch for one call.  Calls that may run concurrently on different streams (the pipelined serving loop:
        encode of batch i+1 under the decode of batch i) must use different tags."""
        n = C.c_size_t()
        _l.check(_l.load().asr_workspace_bytes(self.handle, B, max(T, 7), max(Ldec, 1), C.byref(n)), "workspace_bytes")
        return _l.workspace(n.value, self.device, tag)

    def _encode_impl(self, spectrum: torch.Tensor, enc_lens: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, ws_tag: str = "model") -> torch.Tensor:
        """Transformer.input_layer + Encoder.forward: (B,1,F,T) fp32 -> (B,T',D) fp32."""
        if spectrum.dim() != 4 or spectrum.shape[1] != 1 or spectrum.shape[2] != self.cfg.input_dim:
            raise RuntimeError(f"expected spectrum (B,1,{self.cfg.input_dim},T), got {tuple(spectrum.shape)}")
        spectrum = spectrum.to(torch.float32).contiguous()
        B, _, _, T = spectrum.shape
        Tp = conv_len(conv_len(T))
        if out is None:
            out = torch.empty(B, Tp, self.cfg.embedding_dim, dtype=torch.float32, device=spectrum.device)
        ws = self._ws(B, T, self.cfg.decoder_seq_len, ws_tag)
        lens = None if enc_lens is None else enc_lens.to(device=spectrum.device, dtype=torch.int32).contiguous()
        _l.check(_l.load().asr_encode(self.handle, _l.ptr(spectrum), B, T, _l.ptr(lens), _l.ptr(ws), ws.numel(),
                                      _l.ptr(out), _l.stream()), "asr_encode")
        return out

    def _encoder_forward_impl(self, z_f16: torch.Tensor, enc_lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, Tp, _ = z_f16.shape
        out = torch.empty(B, Tp, self.cfg.embedding_dim, dtype=torch.float32, device=z_f16.device)
        ws = self._ws(B, 4 * Tp + 3, self.cfg.decoder_seq_len)
        lens = None if enc_lens is None else enc_lens.to(device=z_f16.device, dtype=torch.int32).contiguous()
        _l.check(_l.load().asr_encoder_forward(self.handle, _l.ptr(z_f16), B, Tp, _l.ptr(lens), _l.ptr(ws),
                                               ws.numel(), _l.ptr(out), _l.stream()), "asr_encoder_forward")
        return out

    def _decoder_forward_impl(self, enc_out: torch.Tensor, text: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
        B, Tp, _ = enc_out.shape
        L = text.shape[1]
        enc_out = enc_out.to(torch.float32).contiguous()
        text = text.to(torch.int32).contiguous()
        valid = valid.to(torch.uint8).contiguous()
        logits = torch.empty(B, L, self.cfg.vocab_size, dtype=torch.float32, device=enc_out.device)
        ws = self._ws(B, 4 * Tp + 3, L)
        _l.check(_l.load().asr_decoder_forward(self.handle, _l.ptr(enc_out), B, Tp, _l.ptr(text), _l.ptr(valid), L,
                                               _l.ptr(ws), ws.numel(), _l.ptr(logits), _l.stream()),
                 "asr_decoder_forward")
        return logits

    def _decode_greedy_impl(self, enc_out: torch.Tensor, max_len: Optional[int] = None, stop_at_eos: bool = False,
                      first_tokens: Optional[torch.Tensor] = None, want_logits: bool = False,
                      tokens_out: Optional[torch.Tensor] = None, n_tokens_out: Optional[torch.Tensor] = None,
                      enc_lens: Optional[torch.Tensor] = None, ws_tag: str = "model", phase: str = "both"):
        """``phase``: "both" (asr_decode_greedy), "prepare" (cross K/V + state init, returns the call context) or a
        context returned by "prepare" (runs the decode loop on it: the two halves may sit on different streams)."""
        if isinstance(phase, dict):
            ctx = phase
            _l.check(_l.load().asr_decode_run(*ctx["args"], _l.stream()), "asr_decode_run")
            return ctx["tokens"], ctx["n_tok"], ctx["step_logits"]
        B, Tp, _ = enc_out.shape
        L = int(max_len or self.cfg.decoder_seq_len)
        enc_out = enc_out.to(torch.float32).contiguous()
        dev = enc_out.device
        tokens = tokens_out if tokens_out is not None else torch.empty(B, L + 1, dtype=torch.int32, device=dev)
        n_tok = n_tokens_out if n_tokens_out is not None else torch.empty(B, dtype=torch.int32, device=dev)
        step_logits = torch.empty(B, L, self.cfg.vocab_size, dtype=torch.float32, device=dev) if want_logits else None
        first = None if first_tokens is None else first_tokens.to(device=dev, dtype=torch.int32).contiguous()
        lens = None if enc_lens is None else enc_lens.to(device=dev, dtype=torch.int32).contiguous()
        ws = self._ws(B, 4 * Tp + 3, L, ws_tag)
        args = (self.handle, _l.ptr(enc_out), B, Tp, L, int(bool(stop_at_eos)), _l.ptr(first), _l.ptr(lens), _l.ptr(ws),
                ws.numel(), _l.ptr(tokens), _l.ptr(n_tok), _l.ptr(step_logits))
        if phase == "prepare":
            _l.check(_l.load().asr_decode_prepare(*args, _l.stream()), "asr_decode_prepare")
            return {"args": args, "tokens": tokens, "n_tok": n_tok, "step_logits": step_logits,
                    "keep": (enc_out, first, lens, ws)}
        _l.check(_l.load().asr_decode_greedy(*args, _l.stream()), "asr_decode_greedy")
        return tokens, n_tok, step_logits

    def _decode_beam_impl(self, enc_out: torch.Tensor, beam: int, max_len: Optional[int] = None):
        """Beam search on the KV-cached decode step (asr_decode_beam): enc_out (B,T',D) -> tokens (B, beam, L+1) int32
        best first, scores (B, beam) fp32 (sum of token log-probabilities, no length normalisation)."""
        B, Tp, _ = enc_out.shape
        beam = int(beam)
        L = int(max_len or self.cfg.decoder_seq_len)
        dev = enc_out.device
        tokens = torch.empty(B, beam, L + 1, dtype=torch.int32, device=dev)
        scores = torch.empty(B, beam, dtype=torch.float32, device=dev)
        if B == 0:
            return tokens, scores
        enc_rep = enc_out.to(torch.float32).repeat_interleave(beam, 0).contiguous()   # hypotheses are batch rows
        n = C.c_size_t()
        _l.check(_l.load().asr_beam_workspace_bytes(self.handle, B, beam, 4 * Tp + 3, L, C.byref(n)),
                 "asr_beam_workspace_bytes")
        ws = _l.workspace(n.value, self.device, "beam")
        _l.check(_l.load().asr_decode_beam(self.handle, _l.ptr(enc_rep), B, beam, Tp, L, _l.ptr(ws), ws.numel(),
                                           _l.ptr(tokens), _l.ptr(scores), _l.stream()), "asr_decode_beam")
        return tokens, scores
