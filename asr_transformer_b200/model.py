"""Drop-in replacements for reference ``modules/Transformer/model.py``.

Class names, constructor signatures, sub-module names (hence ``state_dict`` keys, including the dead parameters
``input_encoding`` and ``EncoderLayer._norm_in``) and the results of ``forward`` / ``evaluate``
match the reference; the arithmetic runs in sm_100a kernels behind the C-ABI.  ``Transformer.greedy_decode`` is
the additive, batched API (device KV cache, on-GPU argmax/EOS, no per-token host sync).
"""
from __future__ import annotations

import os

from typing import Optional, Tuple

import torch
from torch import nn

from . import lib as _l
from .engine import Engine, conv_len
from .layers import MHA, FeedForward, LayerNorm, TrainablePositionalEncoding, _require_eval


class EncoderLayer(nn.Module):
    """reference model.py:9-25 (pre-LN; ``_norm_in`` exists but is never applied)."""

    def __init__(self, emb_dim, num_heads, ff_dim, dropout):
        super().__init__()
        self._norm_in = LayerNorm(emb_dim)
        self._attention = MHA(num_heads, emb_dim, dropout)
        self._norm1 = LayerNorm(emb_dim)
        self._feedforward = FeedForward(emb_dim, ff_dim, dropout)
        self._norm2 = LayerNorm(emb_dim)

    def forward(self, x):
        # sub-module drop-in built from the operator entry points; Encoder.forward uses the fused path instead
        x = self._attention(self._norm1(x)) + x
        return self._feedforward(self._norm2(x)) + x


class Encoder(nn.Module):
    """reference model.py:28-52."""

    def __init__(self, seq_len, emb_dim, input_dim, num_layers, num_heads, ff_dim, dropout=0.1):
        super().__init__()
        self._lin_in = nn.Linear(input_dim, emb_dim)
        self._norm_out = LayerNorm(emb_dim)
        self._pe = TrainablePositionalEncoding(seq_len, emb_dim)
        self._layers = nn.ModuleList([EncoderLayer(emb_dim, num_heads, ff_dim, dropout) for _ in range(num_layers)])
        self._engine = Engine()

    def forward(self, x):
        """x: conv features (B, 64, F', T') fp32, as produced by ``Transformer.input_layer`` -> (B, T', D)."""
        _require_eval(self)
        eng = self._engine.sync(self, encoder=self)
        B, Cc, Fp, Tp = x.shape
        # NCHW -> the kernel's (B, T', f*64 + c) fp16 feature layout (replaces view/transpose/contiguous, model.py:43-45);
        # rows are [hi | lo] pairs when the library runs with split operands (include/asr_b200.h: asr_split_operands)
        zf = x.permute(0, 3, 2, 1).reshape(B, Tp, Fp * Cc).to(torch.float32)
        z = zf.clamp(-65504.0, 65504.0).to(torch.float16)
        if _l.load().asr_split_operands():
            z = torch.cat([z, (zf - z.float()).to(torch.float16)], -1)
        return eng.encoder_forward(z.contiguous())


class DecoderLayer(nn.Module):
    """reference model.py:55-75."""

    def __init__(self, emb_dim, num_heads, ff_dim, dropout):
        super().__init__()
        self._mask_attention = MHA(num_heads, emb_dim, dropout)
        self._norm1 = LayerNorm(emb_dim)
        self._cross_attention = MHA(num_heads, emb_dim, dropout)
        self._norm2 = LayerNorm(emb_dim)
        self._feedforward = FeedForward(emb_dim, ff_dim, dropout)
        self._norm3 = LayerNorm(emb_dim)

    def forward(self, x, mask, enc_x):
        x = self._mask_attention(self._norm1(x), attention_mask=mask) + x
        x = self._cross_attention(self._norm2(x), enc_x=enc_x) + x
        return self._feedforward(self._norm3(x)) + x


class Decoder(nn.Module):
    """reference model.py:78-151."""

    def __init__(self, vocab_size, seq_len, emb_dim, num_layers, num_heads, ff_dim, eos_token_id, dropout=0.1,
                 pad_token_id=0):
        super().__init__()
        self._seq_len = seq_len
        self._eos_token_id = eos_token_id
        self._embedding = nn.Embedding(vocab_size, emb_dim, padding_idx=pad_token_id)
        self._pe = TrainablePositionalEncoding(seq_len, emb_dim)
        self._dropout = nn.Dropout(dropout)
        self._layers = nn.ModuleList([DecoderLayer(emb_dim, num_heads, ff_dim, dropout) for _ in range(num_layers)])
        self._norm_layer = LayerNorm(emb_dim)
        self._classifier = nn.Linear(emb_dim, vocab_size, bias=False)
        self._engine = Engine()

    def _eng(self) -> Engine:
        return self._engine.sync(self, decoder=self)

    def forward(self, x, mask, enc_x):
        """Teacher-forced logits (B, L, V): mask = pad_k | pad_q | causal, final LayerNorm + classifier
        (reference model.py:104-123)."""
        _require_eval(self)
        return self._eng().decoder_forward(enc_x, x, mask.ge(1))

    def evaluate(self, x, enc_x):
        """Greedy search with the reference's return contract (model.py:125-151): returns
        (decoder_input of the LAST sample (1, L+1) int64, probs list).  No final LayerNorm, exactly L steps, no
        stop at EOS; ``probs`` gets one ``prob[:, :-1].squeeze()`` entry per EOS emission and one at step L."""
        _require_eval(self)
        return _evaluate(self._eng(), x, enc_x, self._seq_len, self._eos_token_id)


def _evaluate(eng: Engine, x: torch.Tensor, enc_x: torch.Tensor, seq_len: int, eos: int):
    if x.dim() != 2 or x.shape[1] != 1:
        raise RuntimeError("asr_b200 evaluate(): expected the start tokens as a (B, 1) tensor (train.py:70)")
    B = x.shape[0]
    tokens, _, step_logits = eng.decode_greedy(enc_x, seq_len, stop_at_eos=False, first_tokens=x[:, 0],
                                               want_logits=True)
    tokens = tokens.long()
    probs = []
    emitted = tokens[:, 1:].cpu()   # token chosen at step i is tokens[:, i], i = 1..L
    for b in range(B):
        for i in range(1, seq_len + 1):
            if int(emitted[b, i - 1]) == eos or i == seq_len:
                # reference: prob is (1, i, V) at step i; prob[:, :-1].squeeze()
                probs.append(step_logits[b:b + 1, :i - 1].squeeze())
    return tokens[B - 1:B], probs


class ConvFrontEnd(nn.Sequential):
    """``Transformer.input_layer`` (reference model.py:168-171): Conv2d(1,64,3,2)+ReLU+Conv2d(64,64,3,2)+ReLU."""

    def forward(self, spectrum):
        L = _l.load()
        from .engine import pack_conv1, pack_conv2_fragments, _f32
        spectrum = spectrum.to(torch.float32).contiguous()
        B, _, F, T = spectrum.shape
        Fp, Tp = conv_len(conv_len(F)), conv_len(conv_len(T))
        dev = spectrum.device
        w1, b1 = pack_conv1(self[0].weight), _f32(self[0].bias)
        w2, b2 = pack_conv2_fragments(self[2].weight), _f32(self[2].bias)
        sp = 2 if L.asr_split_operands() else 1
        z = torch.empty(B, Tp, sp * Fp * 64, dtype=torch.float16, device=dev)
        ws = _l.workspace(L.asr_conv_workspace_bytes(B, F, T), dev, "op")
        _l.check(L.asr_conv_frontend(_l.ptr(spectrum), _l.ptr(w1), _l.ptr(b1), _l.ptr(w2), _l.ptr(b2), B, F, T,
                                     _l.ptr(ws), ws.numel(), _l.ptr(z), _l.stream()), "asr_conv_frontend")
        # kernel layout (B, T', f*64+c) [hi | lo halves] -> the reference's NCHW (B, 64, F', T')
        zf = z[..., :Fp * 64].float() + (z[..., Fp * 64:].float() if sp == 2 else 0.0)
        return zf.view(B, Tp, Fp, 64).permute(0, 3, 2, 1).contiguous()


class Transformer(nn.Module):
    """reference model.py:154-206."""

    def __init__(self, vocab_size, input_dim, embedding_dim, decoder_seq_len, encoder_seq_len, encoder_num_layers,
                 decoder_num_layers, num_heads, ff_dim, dropout=0.1, pad_token_id=4, eos_token_id=2,
                 bos_token_id=1):
        super().__init__()
        self._input_dim = input_dim
        self._bos_token_id = bos_token_id
        self.input_layer = ConvFrontEnd(nn.Conv2d(1, 64, 3, stride=2), nn.ReLU(),
                                        nn.Conv2d(64, 64, 3, stride=2), nn.ReLU())
        feat = conv_len(conv_len(input_dim)) * 64
        self.input_encoding = nn.Linear(feat, embedding_dim)     # dead parameter kept for state_dict parity (Q9)
        self.encoder = Encoder(seq_len=encoder_seq_len, input_dim=feat, emb_dim=embedding_dim,
                               num_layers=encoder_num_layers, num_heads=num_heads, ff_dim=ff_dim, dropout=dropout)
        self.decoder = Decoder(vocab_size=vocab_size, seq_len=decoder_seq_len, emb_dim=embedding_dim,
                               num_layers=decoder_num_layers, num_heads=num_heads, ff_dim=ff_dim,
                               eos_token_id=eos_token_id, dropout=dropout, pad_token_id=pad_token_id)
        self._engine = Engine()

    def _eng(self) -> Engine:
        return self._engine.sync(self, input_layer=self.input_layer, encoder=self.encoder, decoder=self.decoder,
                                 bos=self._bos_token_id)

    # ---- reference API ---------------------------------------------------------------------------------
    def forward(self, spectrum, text, mask):
        """(B,1,F,T), (B,L) tokens, (B,L) mask (>=1 real) -> logits (B,L,V) fp32 (reference model.py:194-198)."""
        _require_eval(self)
        eng = self._eng()
        return eng.decoder_forward(eng.encode(spectrum), text, mask.ge(1))

    def evaluate(self, spectrum, text):
        """Reference greedy search contract (model.py:201-206): (tokens of the LAST sample (1,L+1) int64, probs)."""
        _require_eval(self)
        eng = self._eng()
        return _evaluate(eng, text, eng.encode(spectrum), self.decoder._seq_len, self.decoder._eos_token_id)

    # ---- additive API ----------------------------------------------------------------------------------
    @staticmethod
    def encoder_lengths(lengths: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Input frames per utterance -> encoder frames after the two stride-2 valid convolutions (model.py:168-171)."""
        if lengths is None:
            return None
        l1 = torch.div(lengths - 3, 2, rounding_mode="floor") + 1
        return (torch.div(l1 - 3, 2, rounding_mode="floor") + 1).clamp_min(0)

    def encode(self, spectrum, lengths: Optional[torch.Tensor] = None):
        """Front-end + encoder. ``lengths`` (input frames per utterance) switches on the key-padding mask that the
        reference root encoder lacks; leave None for reference parity."""
        _require_eval(self)
        return self._eng().encode(spectrum, self.encoder_lengths(lengths))

    def greedy_decode(self, spectrum, lengths: Optional[torch.Tensor] = None, max_len: Optional[int] = None,
                      stop_at_eos: bool = False, return_logits: bool = False):
        """Batched greedy ASR: (B,1,F,T) -> tokens (B, L+1) int32 (column 0 = BOS), n_tokens (B,) int32
        [, step_logits (B,L,V)].  stop_at_eos=False decodes exactly L steps like the reference.  ``lengths`` (input
        frames per utterance) masks the zero padding end to end: encoder self attention AND the decoder's cross
        attention ignore the padded frames, so a padded utterance decodes like the unpadded one."""
        _require_eval(self)
        enc = self.encode(spectrum, lengths)
        tokens, n_tok, step_logits = self._eng().decode_greedy(enc, max_len, stop_at_eos, None, return_logits,
                                                               enc_lens=self.encoder_lengths(lengths))
        return (tokens, n_tok, step_logits) if return_logits else (tokens, n_tok)

    def transcribe(self, audio_batches, spectrogram, detokenizer, frames: Optional[int] = None,
                   max_len: Optional[int] = None):
        """Raw audio to text through the whole pipeline (the callers on either side of the hot path, SURVEY.md 8f rows
        2-3): every item of ``audio_batches`` is (B, N) fp32 audio (host or device); ``spectrogram`` a
        ``frontend.Spectrogram`` (the reference's dataset.py:34-35 configuration), ``frames`` the frame count the
        spectrograms are zero-padded / cut to (dataset.py:53-55; default: each batch's own length), ``detokenizer`` a
        ``text.Detokenizer``.  Spectrograms are computed on the device and fed to ``greedy_decode_batches`` (decode
        stops at EOS); yields one list of strings per batch."""
        dev = next(self.parameters()).device

        def specs():
            for a in audio_batches:
                yield spectrogram(a.to(dev, non_blocking=True), frames_out=frames)

        for tokens, n_tok in self.greedy_decode_batches(specs(), max_len=max_len, stop_at_eos=True):
            yield detokenizer.decode_batch(tokens, n_tok)

    def beam_search(self, spectrum, beam: int = 4, lengths: Optional[torch.Tensor] = None,
                    max_len: Optional[int] = None):
        """Beam search (the reference's README TODO; semantics in ``include/asr_b200.h: asr_decode_beam``):
        (B,1,F,T) -> tokens (B, beam, L+1) int32 with the best hypothesis first, scores (B, beam) fp32 = sum of the
        token log-probabilities.  ``beam=1`` equals ``greedy_decode(stop_at_eos=True)``."""
        _require_eval(self)
        if not 1 <= int(beam) <= 16:
            raise ValueError("beam must be in 1..16")
        enc = self.encode(spectrum, lengths)
        return self._eng().decode_beam(enc, beam, max_len)

    def greedy_decode_batches(self, batches, max_len: Optional[int] = None, stop_at_eos: bool = False,
                              gather=None, to_host: bool = True, coalesce: Optional[int] = None):
        """Pipelined greedy ASR over an iterable of batches ((B,1,F,T) fp32; HOST tensors, ideally pinned, or tensors
        already on the device; an item may also be a pair (batch, lengths) to mask the zero padding of each utterance).
        Four streams keep every engine busy: the upload of group i+1, the ENCODER of group i+1 (it runs on the SMs the
        decoder leaves idle: the cluster decoder occupies num_heads x ceil(B / group) SMs and
        is latency-bound, so the two overlap), the decoder of group i (high priority), and the download of group i-1's
        transcripts.  Yields (tokens (B,L+1) int32, n_tokens (B,) int32) per input batch, in order: CPU tensors
        (``to_host``) or device tensors.  ``gather`` (optional callable (tokens, n_tokens) -> (tokens, n_tokens)) runs
        on the device before the download, e.g. ``parallel.gather_tokens`` for the multi-GPU transcript gather.

        ``coalesce``: consecutive same-shaped batches are decoded as one group of up to ``coalesce`` batches (one
        front-end/encoder pass and ONE decode launch per group).  A decode step costs a fixed weight stream and a fixed
        chain of ~70 micro-phases per cluster of CTAs however many utterances share it, so 8 utterances per cluster
        (256 per launch on a 148-SM part) decode in 2.4x the time of 2 per cluster: the default (None) fills launches up
        to 256 utterances (at most 4 batches).  Utterances are independent; the group size only changes how the attention
        keys are dealt to the warps (fp32 summation order), so the tokens equal those of per-batch ``greedy_decode``
        except at argmax near-ties (``coalesce=1`` is bit-identical to the per-batch call)."""
        _require_eval(self)
        from collections import deque
        dev = next(self.parameters()).device
        eng = self._eng()
        streams = self.__dict__.setdefault("_side_streams", {})   # persistent: the caching allocator pools blocks per
        if dev not in streams:                                     # stream, fresh streams would mean fresh cudaMallocs
            streams[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev),
                            torch.cuda.Stream(dev, priority=-1))
        up_s, down_s, enc_s, dec_s = streams[dev]
        caller = torch.cuda.current_stream(dev)
        for st in (up_s, enc_s, dec_s):
            st.wait_stream(caller)                                 # inputs produced on the caller's stream
        up_s.wait_stream(enc_s)      # the staging buffers and decode workspaces are persistent: a previous call that was
        enc_s.wait_stream(dec_s)     # abandoned half way may still have work in flight on them
        staging = self.__dict__.setdefault("_pinned_staging", {})   # pinned D2H buffers live with the model:
        # cudaHostAlloc costs milliseconds, so they are allocated once per shape and reused by every call

        # Both caches hold ONE grow-only flat buffer per slot, viewed per shape: a serving loop with varying batch sizes
        # or frame counts (length-bucketed batches) keeps a bounded set of buffers instead of one per shape.  A buffer
        # that is outgrown may still be in use by work already queued: the device one is handed back to the allocator
        # only after every side stream has passed this point (record_stream), the pinned one is kept alive (growth is
        # geometric, so only a handful are ever retired).
        retired = self.__dict__.setdefault("_retired_staging", [])

        def numel(shape):
            n = 1
            for d in shape:
                n *= int(d)
            return n

        def pinned(t, slot):
            key = (slot, t.dtype)
            n = t.numel()
            flat = staging.get(key)
            if flat is None or flat.numel() < n:
                if flat is not None:
                    retired.append(flat)
                flat = torch.empty(max(n, int(1.5 * flat.numel()) if flat is not None else n), dtype=t.dtype,
                                   pin_memory=True)
                staging[key] = flat
            return flat[:n].view(t.shape)

        bufs = self.__dict__.setdefault("_pipe_bufs", {})

        def stage_buf(kind, shape):
            key = (str(dev),) + kind
            n = numel(shape)
            flat = bufs.get(key)
            if flat is None or flat.numel() < n:
                if flat is not None:
                    for st in (up_s, down_s, enc_s, dec_s):
                        flat.record_stream(st)
                flat = torch.empty(max(n, int(1.5 * flat.numel()) if flat is not None else n), dtype=torch.float32,
                                   device=dev)
                bufs[key] = flat
            return flat[:n].view(tuple(shape))

        def groups():
            """Consecutive batches of one shape / placement / masking, up to `coalesce` per group.  An item is a
            spectrogram batch or a pair (spectrogram batch, lengths): input frames per utterance, which switch on the
            key-padding masks end to end like ``greedy_decode(lengths=...)``."""
            cur, lens, limit = [], [], 1
            for item in batches:
                x, ln = item if isinstance(item, (tuple, list)) else (item, None)
                if cur and (x.shape != cur[0].shape or x.is_cuda != cur[0].is_cuda or (ln is None) != (lens[0] is None)
                            or len(cur) >= limit):
                    yield cur, lens
                    cur, lens = [], []
                if not cur:
                    limit = coalesce if coalesce else max(1, min(4, 256 // max(1, x.shape[0])))
                cur.append(x)
                lens.append(ln)
            if cur:
                yield cur, lens

        def stage_in(group):
            """H2D (if needed) on the upload stream, then the front-end + encoder on the encoder stream."""
            if group is None:
                return None
            group, lens = group
            sizes = [int(x.shape[0]) for x in group]
            lens_all = None if lens[0] is None else torch.cat([torch.as_tensor(l).reshape(-1) for l in lens])
            # Device staging (input batch of the group, encoder output) is persistent per pipeline slot: the caching
            # allocator would otherwise have to serve ~70 MB per group across three streams, and a cudaMalloc in the
            # middle of the loop serialises every stream.  Slot k & 1 is reused by group k + 2: its encoder pass
            # (encoder stream) must have finished before the upload stream overwrites the input buffer.
            slot = n_in[0] & 1
            single = len(group) == 1 and group[0].is_cuda
            x = group[0] if single else stage_buf(("x", slot), (sum(sizes),) + tuple(group[0].shape[1:]))
            if not group[0].is_cuda:
                with torch.cuda.stream(up_s):
                    if x_free[slot] is not None:
                        up_s.wait_event(x_free[slot])
                    o = 0
                    for xb in group:
                        x[o:o + xb.shape[0]].copy_(xb, non_blocking=True)
                        o += xb.shape[0]
                    ev = torch.cuda.Event()
                    ev.record(up_s)
                enc_s.wait_event(ev)
            else:
                for xb in group:
                    xb.record_stream(enc_s)
            with torch.cuda.stream(enc_s):
                if ptrace is not None:
                    ptrace.append({"enc0": tev(enc_s)})
                if group[0].is_cuda and not single:
                    o = 0
                    for xb in group:
                        x[o:o + xb.shape[0]].copy_(xb, non_blocking=True)
                        o += xb.shape[0]
                enc_lens = None if lens_all is None else self.encoder_lengths(lens_all.to(device=dev, dtype=torch.int32))
                Tp = conv_len(conv_len(int(x.shape[-1])))
                if n_in[0] == 0:
                    # size the OTHER slot's buffers like this group's right away: the first group pays every
                    # allocation of the loop, no cudaMalloc (a device-wide synchronisation) happens later
                    L_ws = int(max_len or eng.cfg.decoder_seq_len)
                    eng._ws(sum(sizes), 4 * Tp + 3, L_ws, "pipe_dec1")
                    if not single:
                        stage_buf(("x", 1), tuple(x.shape))
                    stage_buf(("enc", 1), (sum(sizes), Tp, eng.cfg.embedding_dim))
                enc = eng.encode(x, enc_lens, out=stage_buf(("enc", slot), (sum(sizes), Tp, eng.cfg.embedding_dim)),
                                 ws_tag="pipe_enc")
                x_free[slot] = torch.cuda.Event()
                x_free[slot].record(enc_s)
                if ptrace is not None:
                    ptrace[-1]["enc1"] = tev(enc_s)
                # cross-attention K/V + decoder state of this group: also under the previous group's decode loop.  Two
                # decode workspaces alternate; a slot is reused only after the decode that last ran on it has finished.
                if ws_done[slot] is not None:
                    enc_s.wait_event(ws_done[slot])
                ctx = eng.decode_greedy(enc, max_len, stop_at_eos, ws_tag="pipe_dec%d" % slot, phase="prepare",
                                        enc_lens=enc_lens)
                ctx["slot"] = slot
                ctx["sizes"] = sizes
                n_in[0] += 1
                ev = torch.cuda.Event()
                ev.record(enc_s)
                if ptrace is not None:
                    ptrace[-1]["prep1"] = tev(enc_s)
                    ctx["trace"] = ptrace[-1]
            return ctx, ev

        def emit(a, b, sizes):
            o = 0
            for sz in sizes:
                t, n = a[o:o + sz], b[o:o + sz]
                o += sz
                yield (t.clone(), n.clone()) if to_host else (t, n)

        n_in = [0]
        ws_done = [None, None]
        x_free = [None, None]
        # ASR_B200_PIPE_TRACE=1 (diagnostic): timed events around every group's encoder + prepare and decode launch; the
        # timeline is printed to stderr when the loop ends (tools/prof_pipeline.py)
        ptrace = [] if os.environ.get("ASR_B200_PIPE_TRACE") else None

        def tev(stream):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            return e
        it = groups()
        nxt = stage_in(next(it, None))
        pending = deque()
        step = 0
        while nxt is not None:
            ctx, ev = nxt
            sizes = ctx["sizes"]
            dec_s.wait_event(ev)
            for t in (ctx["keep"][0], ctx["tokens"], ctx["n_tok"]):
                t.record_stream(dec_s)
            with torch.cuda.stream(dec_s):
                if ptrace is not None:
                    ctx["trace"]["dec0"] = tev(dec_s)
                tokens, n_tok, _ = eng.decode_greedy(None, phase=ctx)
                done = torch.cuda.Event()
                done.record(dec_s)
                if ptrace is not None:
                    ctx["trace"]["dec1"] = tev(dec_s)
            ws_done[ctx["slot"]] = done
            nxt = stage_in(next(it, None))            # next group: upload + encoder under this group's decode
            # transcript side (its own stream, so that neither the multi-GPU gather - a rendezvous of all ranks - nor
            # the download ever sits between two decode launches)
            with torch.cuda.stream(down_s):
                down_s.wait_event(done)
                tokens.record_stream(down_s)
                n_tok.record_stream(down_s)
                if gather is not None:                 # per input batch: every rank contributes its slice of batch j
                    o, tl, nl = 0, [], []
                    for sz in sizes:
                        t, n = gather(tokens[o:o + sz], n_tok[o:o + sz])
                        tl.append(t)
                        nl.append(n)
                        o += sz
                    sizes = [int(t.shape[0]) for t in tl]
                    tokens, n_tok = (tl[0], nl[0]) if len(tl) == 1 else (torch.cat(tl, 0), torch.cat(nl, 0))
                if to_host:
                    th, nh = pinned(tokens, (step % 3, 0)), pinned(n_tok, (step % 3, 1))
                    th.copy_(tokens, non_blocking=True)
                    nh.copy_(n_tok, non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(down_s)
            if to_host:
                pending.append((th, nh, fin, sizes))
            else:
                tokens.record_stream(caller)
                n_tok.record_stream(caller)
                pending.append((tokens, n_tok, fin, sizes))
            step += 1
            if len(pending) > 1:
                a, b, e, sz = pending.popleft()
                if to_host:
                    e.synchronize()
                else:
                    caller.wait_event(e)
                yield from emit(a, b, sz)
        while pending:
            a, b, e, sz = pending.popleft()
            if to_host:
                e.synchronize()
            else:
                caller.wait_event(e)
            yield from emit(a, b, sz)
        if ptrace:
            import sys
            torch.cuda.synchronize(dev)
            t0 = ptrace[0]["enc0"]
            print("group | encoder start end | prepare end | decode start end   (ms since the first encoder start)", file=sys.stderr)
            for i, r in enumerate(ptrace):
                if "dec1" in r:
                    print("%5d | %8.3f %8.3f | %8.3f | %8.3f %8.3f" % ((i,) + tuple(t0.elapsed_time(r[k]) for k in (
                        "enc0", "enc1", "prep1", "dec0", "dec1"))), file=sys.stderr)
