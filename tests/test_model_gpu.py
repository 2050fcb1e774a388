"""Model-level parity on the GPU through the drop-in modules / C-ABI:
golden fixtures produced by the unmodified reference (T0, C1) and the CPU oracle at BASELINE sizes (C2 ...)."""
import os

import pytest
import torch

from oracle import speech_transformer as O
from tests.util import TAU, TOL_FP32, assert_close, build_model, cpu_state, golden, state_checksum

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def t0():
    cfg = O.CONFIGS["T0"]
    fx = golden("model_T0.pt")
    m = build_model(cfg, DEV)
    assert state_checksum(cpu_state(m)) == fx["state_checksum"]
    spec = O.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=1).to(DEV)
    return cfg, fx, m, spec


def test_frontend_and_encoder_T0(t0):
    cfg, fx, m, spec = t0
    conv = m.input_layer(spec)
    assert_close(conv, fx["conv_out"], 6e-2, 4e-3, "input_layer")
    enc = m.encode(spec)
    assert_close(enc, fx["enc_out"], what="enc_out")
    # Encoder.forward drop-in on reference-layout conv features
    assert_close(m.encoder(fx["conv_out"].to(DEV)), fx["enc_out"], what="Encoder.forward")


def test_forward_logits_T0(t0):
    cfg, fx, m, spec = t0
    logits = m(spec, fx["text"].to(DEV), fx["mask"].to(DEV))
    assert logits.dtype == torch.float32 and logits.shape == fx["forward_logits"].shape
    assert_close(logits, fx["forward_logits"], what="Transformer.forward logits")
    # Decoder.forward drop-in on the reference's encoder output
    out = m.decoder(fx["text"].to(DEV), fx["mask"].to(DEV), fx["enc_out"].to(DEV))
    assert_close(out, fx["forward_logits"], what="Decoder.forward logits")


def check_tokens(ref_tokens, ref_logits, got_tokens, min_frac=0.99):
    """The north-star bar: greedy tokens identical on >= 99 % of the utterances, every divergence a proven argmax
    near-tie (fp32 reference top1 - top2 margin < TAU at the first differing step)."""
    r = O.compare_tokens(ref_tokens, ref_logits, got_tokens, TAU)
    assert not r["hard"], f"token divergence not explained by an argmax near-tie: {r['hard']}"
    frac = r["identical"] / r["utterances"]
    assert frac >= min_frac, f"only {r['identical']}/{r['utterances']} utterances token-identical (< {min_frac}): {r}"
    return r, frac


@pytest.mark.parametrize("mode", ["cluster", "graph", "eager"])
def test_greedy_tokens_T0(t0, mode, monkeypatch):
    """All three launch modes of asr_decode_greedy (ASR_B200_DECODE) against the reference's tokens and logits."""
    monkeypatch.setenv("ASR_B200_DECODE", mode)
    cfg, fx, m, spec = t0
    tokens, n_tok, logits = m.greedy_decode(spec, return_logits=True)
    torch.cuda.synchronize()
    assert tokens.shape == (cfg.batch, cfg.decoder_seq_len + 1) and tokens.dtype == torch.int32
    r, frac = check_tokens(fx["tokens"], fx["step_logits"], tokens)
    assert frac == 1.0, r
    assert_close(logits, fx["step_logits"], what="step logits (no final LayerNorm)")


def test_evaluate_contract_T0(t0):
    cfg, fx, m, spec = t0
    bos = torch.full((cfg.batch, 1), cfg.bos_token_id, dtype=torch.int32, device=DEV)
    last, probs = m.evaluate(spec, bos)
    assert last.dtype == torch.int64 and tuple(last.shape) == (1, cfg.decoder_seq_len + 1)
    assert torch.equal(last.cpu(), fx["evaluate_tokens_last"])
    assert [tuple(p.shape) for p in probs] == fx["evaluate_probs_shapes"]
    got = torch.tensor([float(p.double().sum()) for p in probs])
    assert torch.allclose(got, fx["evaluate_probs_sum"], rtol=2e-3, atol=0.5)
    # Decoder.evaluate drop-in
    last2, probs2 = m.decoder.evaluate(bos, fx["enc_out"].to(DEV))
    assert torch.equal(last2.cpu(), fx["evaluate_tokens_last"]) and len(probs2) == len(probs)


def test_stop_at_eos_and_lengths(t0):
    cfg, fx, m, spec = t0
    ref = fx["tokens"]
    tokens, n_tok = m.greedy_decode(spec, stop_at_eos=True)
    tokens, n_tok = tokens.cpu().long(), n_tok.cpu()
    for b in range(cfg.batch):
        eos_pos = (ref[b, 1:] == cfg.eos_token_id).nonzero()
        if eos_pos.numel():
            n = int(eos_pos[0]) + 2
            assert int(n_tok[b]) == n
            assert torch.equal(tokens[b, :n], ref[b, :n]) and (tokens[b, n:] == cfg.pad_token_id).all()
        else:
            assert int(n_tok[b]) == cfg.decoder_seq_len + 1 and torch.equal(tokens[b], ref[b])
    # shorter max_len is a prefix of the full decode
    t8, _ = m.greedy_decode(spec, max_len=8)
    assert torch.equal(t8.cpu().long(), ref[:, :9])


def test_batch_invariance_T0(t0):
    """Tokens of an utterance must not depend on the batch it is decoded in."""
    cfg, fx, m, spec = t0
    full, _ = m.greedy_decode(spec)
    for b in range(cfg.batch):
        one, _ = m.greedy_decode(spec[b:b + 1])
        assert torch.equal(one[0], full[b])
    rep = torch.cat([spec, spec, spec], 0)
    t3, _ = m.greedy_decode(rep)
    assert torch.equal(t3[:cfg.batch], full) and torch.equal(t3[2 * cfg.batch:], full)


@pytest.mark.parametrize("batch,gu", [(5, ""), (64, ""), (70, ""), (128, ""), (9, "8"), (200, "")])
def test_cluster_matches_per_kernel_step_C2(batch, gu, monkeypatch):
    """The cluster kernel (head-parallel CTAs, DSMEM all-reduces) against the per-kernel (graph) step for utterance
    groups of 1, 2, 4 and 8 per cluster, incl. stop_at_eos and a ragged last cluster."""
    cfg = O.CONFIGS["C2"]
    m = build_model(cfg, DEV)
    L = 40
    spec = O.structured_spectrum(batch, cfg.frames, cfg.input_dim, seed=23).to(DEV)
    monkeypatch.setenv("ASR_B200_DECODE", "graph")
    tg, ng, lg = m.greedy_decode(spec, max_len=L, return_logits=True)
    sg, nsg = m.greedy_decode(spec, max_len=L, stop_at_eos=True)
    monkeypatch.setenv("ASR_B200_DECODE", "cluster")
    monkeypatch.setenv("ASR_B200_CLUSTER_GU", gu)
    tc, nc, lc = m.greedy_decode(spec, max_len=L, return_logits=True)
    sc, nsc = m.greedy_decode(spec, max_len=L, stop_at_eos=True)
    torch.cuda.synchronize()
    r = O.compare_tokens(tg, lg.cpu(), tc, TAU)
    assert not r["hard"] and r["identical"] >= batch - max(1, batch // 50), r    # two fp32 summation orders of one arithmetic
    same = [b for b in range(batch) if torch.equal(tg[b], tc[b])]
    assert_close(lc[same], lg[same], 5e-3, 2e-4, "cluster vs graph step logits")
    for b in same:
        assert int(nsc[b]) == int(nsg[b]) and torch.equal(sc[b], sg[b])


@pytest.mark.parametrize("batch,gu", [(3, ""), (5, "4"), (9, "8")])
def test_cluster_long_decode_C4(batch, gu, monkeypatch):
    """Long-form config (L = 384, 749 encoder frames): the self attention of the cluster kernel runs over more than one
    super-chunk (256 keys at 2 / 4 utterances per cluster, 128 at 8) and the cross attention over three to six; all
    384 steps against the per-kernel (graph) step."""
    cfg = O.CONFIGS["C4"]
    m = build_model(cfg, DEV)
    spec = O.structured_spectrum(batch, cfg.frames, cfg.input_dim, seed=29).to(DEV)
    monkeypatch.setenv("ASR_B200_DECODE", "graph")
    tg, ng, lg = m.greedy_decode(spec, return_logits=True)
    monkeypatch.setenv("ASR_B200_DECODE", "cluster")
    monkeypatch.setenv("ASR_B200_CLUSTER_GU", gu)
    tc, nc, lc = m.greedy_decode(spec, return_logits=True)
    torch.cuda.synchronize()
    assert tc.shape == (batch, cfg.decoder_seq_len + 1)
    r = O.compare_tokens(tg, lg.cpu(), tc, TAU)
    assert not r["hard"], r
    same = [b for b in range(batch) if torch.equal(tg[b], tc[b])]
    assert len(same) >= batch - 1, r
    assert_close(lc[same], lg[same], 5e-3, 2e-4, "cluster vs graph step logits (L=384)")


def test_empty_batch(t0):
    cfg, fx, m, spec = t0
    tokens, n_tok = m.greedy_decode(spec[:0])
    assert tokens.shape == (0, cfg.decoder_seq_len + 1)


def test_key_padding_truncation_equivalence(t0):
    """Masks-on encoder (the 'next' row): a zero-padded utterance with lengths == the unpadded utterance."""
    cfg, fx, m, spec = t0
    T_short = 123
    padded = spec.clone()
    padded[1, :, :, T_short:] = 0
    lengths = torch.tensor([cfg.frames, T_short, cfg.frames], device=DEV)
    enc = m.encode(padded, lengths)
    alone = m.encode(spec[1:2, :, :, :T_short].contiguous())
    n = alone.shape[1]
    assert_close(enc[1, :n], alone[0], 2e-2, 2e-3, "masked batch row == unpadded utterance")


def test_reference_golden_C1():
    """BASELINE config 1 (repo-default model, batch 8, 10 s, L=128) against the reference's own outputs."""
    cfg = O.CONFIGS["C1"]
    fx = golden("model_C1.pt")
    m = build_model(cfg, DEV)
    assert state_checksum(cpu_state(m)) == fx["state_checksum"]
    spec = O.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=1).to(DEV)
    enc = m.encode(spec)
    assert_close(enc, fx["enc_out"], what="C1 enc_out")
    logits = m(spec, fx["text"].to(DEV), fx["mask"].to(DEV))
    assert_close(logits, fx["forward_logits"], what="C1 forward logits")
    tokens, n_tok, step_logits = m.greedy_decode(spec, return_logits=True)
    r, frac = check_tokens(fx["tokens"], fx["step_logits"], tokens, min_frac=1.0)   # 8/8 against the reference's tokens
    print("C1 greedy:", r)
    ident = [b for b in range(cfg.batch) if torch.equal(tokens[b].cpu().long(), fx["tokens"][b])]
    assert_close(step_logits[ident], fx["step_logits"][ident], what="C1 step logits")


def _greedy_vs_oracle(name, batch, seed=11):
    """One BASELINE size against the CPU oracle (KV-cached restatement, pinned to the reference by the goldens)."""
    cfg = O.CONFIGS[name]
    m = build_model(cfg, DEV)
    sd = cpu_state(m)
    spec = O.structured_spectrum(batch, cfg.frames, cfg.input_dim, seed=seed)
    torch.set_num_threads(os.cpu_count() or 1)
    enc_ref = O.encode(sd, spec)
    tok_ref, logits_ref = O.greedy_kv_cached(sd, enc_ref, cfg)
    enc = m.encode(spec.to(DEV))
    assert_close(enc, enc_ref, 1e-3, 1e-4, what=f"{name} enc_out")     # (split operands: 8e-5 / 1.1e-5 measured)
    tokens, _, step_logits = m.greedy_decode(spec.to(DEV), return_logits=True)
    r = O.compare_tokens(tok_ref, logits_ref, tokens, TAU)
    print(f"{name}@{batch} greedy:", {k: r[k] for k in ("utterances", "identical", "near_tie", "hard", "distinct_rows")})
    assert not r["hard"], f"token divergence not explained by an argmax near-tie: {r['hard']}"
    assert r["distinct_rows"] >= 0.85 * batch      # (the synthetic generator: 229 distinct reference rows of 256)
    # what flips sits at the fp32 reference's own noise floor: margins 1e-4 and below against logits of magnitude 1
    assert all(mg < 1e-4 for _, _, mg in r["near_tie"]), r
    ident = [b for b in range(batch) if torch.equal(tokens[b].cpu().long(), tok_ref[b])]
    d = (step_logits[ident].cpu() - logits_ref[ident]).abs()
    print(f"{name}@{batch} step logits of the identical utterances: max |d| {d.max():.2e}, mean |d| {d.mean():.2e}")
    assert d.max() < 5e-3 and d.mean() < 1e-4
    return r


def test_greedy_vs_oracle_baseline_sizes():
    """The north-star bar on the BASELINE sizes: greedy tokens identical on >= 99 % of the utterances, every divergence a
    proven argmax near-tie.  C2 at 64 / 128 / 256 utterances, C3 (12 encoder layers) at 64, C5 (d_model 512) at 8.
    A single utterance is 0.8 - 1.6 % of one of these batches, and the flips that remain are coin tosses at reference
    margins of 1e-5 (the decoder reads fp16 K/V: step logits carry a mean error of 1.7e-5), so the 99 % bar is asserted on
    the pooled 520 utterances, with a per-size floor of 98 %."""
    res = [_greedy_vs_oracle(n, b) for n, b in (("C2", 64), ("C2", 128), ("C2", 256), ("C3", 64), ("C5", 8))]
    for r in res:
        assert r["identical"] >= 0.98 * r["utterances"], r
    tot, same = sum(r["utterances"] for r in res), sum(r["identical"] for r in res)
    print(f"pooled: {same}/{tot} utterances token-identical ({100.0 * same / tot:.2f} %)")
    assert same >= 0.99 * tot, (same, tot)


def test_layer_dropins_match_oracle_T0(t0):
    """EncoderLayer.forward / DecoderLayer.forward as stand-alone sub-module drop-ins (reference model.py:18-25, 65-75)
    against the oracle's restatement of the same layers, incl. the (S, S) uint8 causal mask of Decoder.evaluate."""
    cfg, fx, m, spec = t0
    sd = cpu_state(m)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 37, cfg.embedding_dim, generator=g)
    enc_ref = O.encoder_layer(sd, "encoder._layers.1", x)
    assert_close(m.encoder._layers[1](x.to(DEV)), enc_ref, 2e-3, 2e-4, "EncoderLayer.forward")
    y = torch.randn(2, 9, cfg.embedding_dim, generator=g)
    mem = torch.randn(2, 37, cfg.embedding_dim, generator=g)
    causal = torch.triu(torch.ones(9, 9, dtype=torch.uint8), diagonal=1)
    dec_ref = O.decoder_layer(sd, "decoder._layers.0", y, causal, mem)
    out = m.decoder._layers[0](y.to(DEV), causal.to(DEV), mem.to(DEV))
    assert_close(out, dec_ref, 2e-3, 2e-4, "DecoderLayer.forward")


def test_masks_on_C5():
    """BASELINE config 5 (d_model 512, 8 heads, FFN 2048) with the key-padding masks ON: (a) the operator oracle, MHA
    with attention_mask (layers.py:22-23), at C5's width; (b) truncation equivalence end to end: a zero-padded utterance
    decoded with its length gives the encoder output and the tokens of the unpadded utterance decoded alone."""
    cfg = O.CONFIGS["C5"]
    m = build_model(cfg, DEV)
    sd = cpu_state(m)
    D = cfg.embedding_dim
    g = torch.Generator().manual_seed(9)
    x, src = torch.randn(3, 21, D, generator=g), torch.randn(3, 50, D, generator=g)
    klen = torch.tensor([50, 17, 33])
    mask = (torch.arange(50)[None, None, :] >= klen[:, None, None]).expand(3, 21, 50)
    ref = O.mha(sd, "decoder._layers.2._cross_attention", x, src, mask)
    out = m.decoder._layers[2]._cross_attention(x.to(DEV), src.to(DEV), mask.to(DEV))
    assert_close(out, ref, 2e-3, 2e-4, "C5 MHA with key-padding mask")
    lens = torch.tensor([cfg.frames, 611, 403, 877])
    spec = O.structured_spectrum(4, cfg.frames, cfg.input_dim, seed=51, lengths=lens)
    enc = m.encode(spec.to(DEV), lens.to(DEV))
    tok, _, lg = m.greedy_decode(spec.to(DEV), lengths=lens.to(DEV), max_len=48, return_logits=True)
    for b in (1, 2, 3):
        n = int(lens[b])
        alone = spec[b:b + 1, :, :, :n].contiguous().to(DEV)
        enc1 = m.encode(alone)
        assert_close(enc[b, :enc1.shape[1]], enc1[0], 1e-3, 1e-4, "C5 masked batch row == unpadded utterance")
        t1, _, lg1 = m.greedy_decode(alone, max_len=48, return_logits=True)
        r = O.compare_tokens(t1.cpu(), lg1.cpu(), tok[b:b + 1].cpu(), TAU)
        assert not r["hard"], r
        assert_close(lg[b, :8], lg1[0, :8], 2e-3, 2e-4, "C5 masked step logits == unpadded")
    # the full-length row equals the unmasked reference-parity decode of that utterance
    ref_tok, ref_lg = O.greedy_kv_cached(sd, O.encode(sd, spec[:1]), cfg, max_len=48)
    r = O.compare_tokens(ref_tok, ref_lg, tok[:1].cpu(), TAU)
    assert not r["hard"] and r["identical"] == 1, r


def test_long_form_encoder_C4():
    """30 s utterances: encoder sequence 749 stresses the flash-attention tiling (6 KV tiles)."""
    cfg = O.CONFIGS["C4"]
    m = build_model(cfg, DEV)
    sd = cpu_state(m)
    spec = O.structured_spectrum(2, cfg.frames, cfg.input_dim, seed=12)
    enc_ref = O.encode(sd, spec)
    assert_close(m.encode(spec.to(DEV)), enc_ref, what="C4 enc_out")
    tok_ref, logits_ref = O.greedy_kv_cached(sd, enc_ref, cfg, max_len=48)
    tokens, _ = m.greedy_decode(spec.to(DEV), max_len=48)
    r, frac = check_tokens(tok_ref, logits_ref, tokens)
    assert frac == 1.0 or not r["hard"]


def test_pipelined_batches_match_single_calls(t0):
    """greedy_decode_batches (H2D / D2H on side streams) returns exactly what per-batch greedy_decode returns."""
    cfg, fx, m, spec = t0
    host = [spec.cpu().pin_memory(), spec.flip(0).cpu().pin_memory(), spec[:2].cpu().pin_memory()]
    outs = list(m.greedy_decode_batches(host))
    assert len(outs) == 3
    for x, (tok, n) in zip(host, outs):
        t_ref, n_ref = m.greedy_decode(x.to(DEV))
        assert not tok.is_cuda and torch.equal(tok, t_ref.cpu()) and torch.equal(n, n_ref.cpu())
    assert list(m.greedy_decode_batches([])) == []


@pytest.mark.parametrize("coalesce", [1, 2, 3])
def test_coalesced_batches_match_single_calls(t0, coalesce):
    """Coalescing consecutive batches into one decode launch (serving throughput) never changes an utterance's tokens:
    host and device inputs, groups cut by the limit, by a shape change and by the end of the stream."""
    cfg, fx, m, spec = t0
    xs = [spec, spec.flip(0), spec.roll(1, 0), spec[:3], spec[:3].flip(0), spec]
    ref = [m.greedy_decode(x.to(DEV)) for x in xs]
    for place in ("host", "device"):
        ins = [x.cpu().pin_memory() if place == "host" else x.to(DEV) for x in xs]
        outs = list(m.greedy_decode_batches(ins, coalesce=coalesce, to_host=(place == "host")))
        assert len(outs) == len(xs)
        for (t_ref, n_ref), (tok, n) in zip(ref, outs):
            assert tok.is_cuda == (place == "device")
            assert torch.equal(tok.cpu(), t_ref.cpu()) and torch.equal(n.cpu(), n_ref.cpu())


def test_coalesced_serving_C2_full_size():
    """BASELINE config 2 through the serving loop exactly as bench.py drives it: four 64-utterance batches share one
    decode launch (256 utterances, 8 per CTA cluster, all 128 steps); a fifth is decoded alone.  The per-batch call
    runs 2 utterances per cluster, which splits the attention keys over the warps differently (another fp32 summation
    order), so the comparison is the usual one: identical tokens, or a first divergence at a proven argmax near-tie."""
    cfg = O.CONFIGS["C2"]
    m = build_model(cfg, DEV)
    xs = [O.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=31 + i) for i in range(5)]
    ref = [m.greedy_decode(x.to(DEV), return_logits=True) for x in xs]
    outs = list(m.greedy_decode_batches([x.pin_memory() for x in xs]))
    assert len(outs) == 5
    for (t_ref, n_ref, lg_ref), (tok, n) in zip(ref, outs):
        r = O.compare_tokens(t_ref.cpu(), lg_ref.cpu(), tok, TAU)
        assert not r["hard"] and r["identical"] >= cfg.batch - 1, r
    # the fifth batch is decoded alone (2 per cluster, like the reference call): bit-exact
    assert torch.equal(outs[4][0], ref[4][0].cpu()) and torch.equal(outs[4][1], ref[4][1].cpu())
    assert len({tuple(r) for r in outs[0][0].tolist()}) >= 0.9 * cfg.batch
    # two batches per launch (4 utterances per cluster)
    outs2 = list(m.greedy_decode_batches([x.pin_memory() for x in xs[:2]], coalesce=2))
    for (t_ref, n_ref, lg_ref), (tok, n) in zip(ref, outs2):
        r = O.compare_tokens(t_ref.cpu(), lg_ref.cpu(), tok, TAU)
        assert not r["hard"] and r["identical"] >= cfg.batch - 1, r


def test_c5_one_launch_of_64_with_masks():
    """BASELINE config 5 as the bench drives it: d_model 512 / 8 heads (clusters of 8 CTAs), 64 mixed-length utterances
    with key-padding masks in ONE decode launch (8 utterances per cluster).  Compared with the same utterances decoded
    16 at a time (2 per cluster: another split of the attention keys over the warps, i.e. another fp32 summation
    order): identical tokens, or a first divergence at a proven argmax near-tie."""
    cfg = O.CONFIGS["C5"]
    m = build_model(cfg, DEV)
    lens = torch.randint(400, 1001, (64,), generator=torch.Generator().manual_seed(5))
    x = O.structured_spectrum(64, 1000, cfg.input_dim, seed=500, lengths=lens).to(DEV)
    ln = lens.to(DEV)
    tok, n = m.greedy_decode(x, lengths=ln)
    same = 0
    for i in range(0, 64, 16):
        t_ref, n_ref, lg_ref = m.greedy_decode(x[i:i + 16], lengths=ln[i:i + 16], return_logits=True)
        r = O.compare_tokens(t_ref.cpu(), lg_ref.cpu(), tok[i:i + 16].cpu(), TAU)
        assert not r["hard"], r
        same += r["identical"]
    assert same >= 62, same
    assert len({tuple(r) for r in tok.cpu().tolist()}) >= 58
    assert torch.equal(tok, m.greedy_decode(x, lengths=ln)[0])       # run-to-run identical


def test_pipelined_batches_with_lengths(t0):
    """(batch, lengths) items switch the key-padding masks on inside the serving loop: same tokens as the per-batch call
    with lengths, for coalesced groups as well, and different from the unmasked decode of the padded input."""
    cfg, fx, m, spec = t0
    lens = torch.tensor([cfg.frames, 131, 160])[:spec.shape[0]]
    padded = spec.clone()
    for b, n in enumerate(lens.tolist()):
        padded[b, :, :, n:] = 0
    items = [(padded.cpu().pin_memory(), lens), (padded.flip(0).cpu().pin_memory(), lens.flip(0)), (padded.cpu(), lens)]
    outs = list(m.greedy_decode_batches(items))
    assert len(outs) == 3
    for (x, ln), (tok, n) in zip(items, outs):
        t_ref, n_ref = m.greedy_decode(x.to(DEV), lengths=ln.to(DEV))
        assert torch.equal(tok, t_ref.cpu()) and torch.equal(n, n_ref.cpu())
    t_unmasked, _ = m.greedy_decode(padded)
    assert not torch.equal(outs[0][0], t_unmasked.cpu())


def test_key_padding_end_to_end_decode(t0):
    """Masks on, end to end: a zero-padded utterance decoded with its length gives the tokens of
    the unpadded utterance decoded alone (encoder self attention and decoder cross attention both ignore the padding)."""
    cfg, fx, m, spec = t0
    T_short = 131
    padded = spec.clone()
    padded[1, :, :, T_short:] = 0
    lengths = torch.tensor([cfg.frames, T_short, cfg.frames], device=DEV)
    tok, _, lg = m.greedy_decode(padded, lengths=lengths, return_logits=True)
    alone, _, lg1 = m.greedy_decode(spec[1:2, :, :, :T_short].contiguous(), return_logits=True)
    full, _ = m.greedy_decode(spec)
    assert torch.equal(tok[0], full[0]) and torch.equal(tok[2], full[2])       # unpadded rows unaffected
    assert_close(lg[1], lg1[0], 2e-2, 2e-3, "masked batch row == unpadded utterance (step logits)")
    r = O.compare_tokens(alone.cpu(), lg1.cpu(), tok[1:2].cpu(), TAU)
    assert not r["hard"], r
    # without the lengths the padded frames leak into the cross attention: the logits differ
    _, _, lg_nomask = m.greedy_decode(padded, return_logits=True)
    assert (lg_nomask[1] - lg1[0]).abs().max() > 1e-3


# ----------------------------------------------------------------------------------------------- beam search (8f rank 4)
def _check_beams(tok_ref, sc_ref, tok, sc, what):
    """Identical hypotheses, or - when fp32 rounding reorders two candidates - scores that agree to 1e-3."""
    tok, sc = tok.cpu().long(), sc.cpu()
    assert tok.shape == tok_ref.shape and sc.shape == sc_ref.shape, what
    fin = torch.isfinite(sc_ref)
    assert torch.equal(fin, torch.isfinite(sc)), what
    assert (sc[fin] - sc_ref[fin]).abs().max().item() < 5e-2, (what, sc, sc_ref)
    same = int((tok == tok_ref).all(-1).sum())
    best_same = int((tok[:, 0] == tok_ref[:, 0]).all(-1).sum())
    return same / max(1, tok_ref.shape[0] * tok_ref.shape[1]), best_same / max(1, tok_ref.shape[0])


@pytest.mark.parametrize("beam", [1, 3, 4])
def test_beam_search_T0(t0, beam):
    """The CUDA beam search against the CPU oracle's definition on the tiny config; beam 1 == greedy that pads at EOS."""
    cfg, fx, m, spec = t0
    sd = cpu_state(m)
    tok_ref, sc_ref = O.beam_search_kv_cached(sd, O.encode(sd, spec.cpu()), cfg, beam)
    tok, sc = m.beam_search(spec, beam=beam)
    frac, best = _check_beams(tok_ref, sc_ref, tok, sc, f"beam {beam}")
    assert best == 1.0 and frac >= 0.9, (frac, best)
    if beam == 1:
        tg, _ = m.greedy_decode(spec, stop_at_eos=True)
        assert torch.equal(tok[:, 0], tg)
    else:   # a wider beam never scores worse than greedy under the same scoring
        t1, s1 = m.beam_search(spec, beam=1)
        assert (sc[:, 0] >= s1[:, 0] - 1e-4).all()
        assert (sc[:, :-1] >= sc[:, 1:]).all()


def test_beam_search_C2_sizes():
    """BASELINE model size (C2 weights, 10 s utterances), 8 utterances x beam 4, 48 steps, against the oracle."""
    cfg = O.CONFIGS["C2"]
    m = build_model(cfg, DEV)
    sd = cpu_state(m)
    spec = O.structured_spectrum(8, cfg.frames, cfg.input_dim, seed=41)
    torch.set_num_threads(os.cpu_count() or 1)
    tok_ref, sc_ref = O.beam_search_kv_cached(sd, O.encode(sd, spec), cfg, 4, max_len=48)
    tok, sc = m.beam_search(spec.to(DEV), beam=4, max_len=48)
    frac, best = _check_beams(tok_ref, sc_ref, tok, sc, "C2 beam 4")
    assert best >= 0.75 and frac >= 0.6, (frac, best)   # near-equal hypotheses may swap places
