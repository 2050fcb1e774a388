"""The CPU oracle against the committed golden fixtures (outputs of the unmodified reference) and, when the
reference tree is importable (build container only), against the live reference."""
import os
import sys

import pytest
import torch

from oracle import speech_transformer as O
from tests.util import build_model, cpu_state, golden, state_checksum

REF = os.environ.get("ASR_REF", "/root/reference")


@pytest.fixture(scope="module")
def t0():
    cfg = O.CONFIGS["T0"]
    fx = golden("model_T0.pt")
    sd = cpu_state(build_model(cfg))
    return cfg, fx, sd


def test_weights_match_reference_seed(t0):
    cfg, fx, sd = t0
    assert state_checksum(sd) == fx["state_checksum"]


def test_conv_len():
    assert O.subsampled_len(1000) == 249 and O.subsampled_len(3000) == 749 and O.subsampled_len(311) == 77
    assert O.conv_len(O.conv_len(80)) == 19 and O.conv_len(O.conv_len(513)) == 127


def test_frontend_encoder_T0(t0):
    cfg, fx, sd = t0
    spec = O.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=1)
    conv = O.frontend(sd, spec)
    assert torch.allclose(conv, fx["conv_out"], atol=1e-5)
    assert torch.allclose(O.encoder(sd, conv), fx["enc_out"], atol=2e-5)


def test_decoder_forward_T0(t0):
    cfg, fx, sd = t0
    out = O.decoder_forward(sd, fx["text"], fx["mask"], fx["enc_out"])
    assert torch.allclose(out, fx["forward_logits"], atol=2e-4)


def test_greedy_T0_both_restatements(t0):
    cfg, fx, sd = t0
    tok, logits = O.greedy_kv_cached(sd, fx["enc_out"], cfg)
    assert torch.equal(tok, fx["tokens"])
    assert torch.allclose(logits[:, :-1], fx["last_probs"], atol=2e-4)
    bos = torch.full((cfg.batch, 1), cfg.bos_token_id, dtype=torch.int32)
    last, probs, rows = O.decoder_evaluate_reference_style(sd, bos, fx["enc_out"], cfg.decoder_seq_len, cfg.eos_token_id)
    assert torch.equal(rows, fx["tokens"])
    assert torch.equal(last, fx["evaluate_tokens_last"]) and last.dtype == torch.int64
    assert [tuple(p.shape) for p in probs] == fx["evaluate_probs_shapes"]
    got = torch.tensor([float(p.double().sum()) for p in probs])
    assert torch.allclose(got, fx["evaluate_probs_sum"], rtol=1e-4, atol=1e-2)


def test_greedy_C1_tokens():
    cfg = O.CONFIGS["C1"]
    fx = golden("model_C1.pt")
    sd = cpu_state(build_model(cfg))
    assert state_checksum(sd) == fx["state_checksum"]
    tok, logits = O.greedy_kv_cached(sd, fx["enc_out"], cfg)
    assert torch.equal(tok, fx["tokens"])
    assert fx["distinct_rows"] == cfg.batch        # structured inputs give distinct rows
    assert torch.allclose(logits[:, :-1], fx["last_probs"], atol=5e-4)


def test_mha_operator_masks():
    fx = golden("ops_mha.pt")
    sd = {"m." + k: v for k, v in fx["state"].items()}
    x, src = fx["x"], fx["src"]
    assert torch.allclose(O.mha(sd, "m", x), fx["self_nomask"], atol=1e-5)
    assert torch.allclose(O.mha(sd, "m", x, src), fx["cross_nomask"], atol=1e-5)
    assert torch.allclose(O.mha(sd, "m", x, src, fx["keypad"]), fx["cross_keypad"], atol=1e-5)
    assert torch.allclose(O.mha(sd, "m", x, None, fx["causal"]), fx["self_causal"], atol=1e-5)
    out = O.mha(sd, "m", x, None, fx["full_rows"])
    assert torch.allclose(out, fx["self_full_rows"], atol=1e-5)
    # fully masked rows -> zeros through nan_to_num -> output equals the out-projection bias
    assert torch.allclose(out[:, 5], fx["state"]["_out_linear.bias"].expand(3, -1), atol=1e-6)
    # truncation equivalence: key-padding mask == running on the unpadded keys
    assert torch.allclose(O.mha(sd, "m", x[1:2], src[1:2, :20]), fx["cross_keypad"][1:2], atol=1e-5)
    ffsd = {"f." + k: v for k, v in fx["ffn_state"].items()}
    assert torch.allclose(O.feed_forward(ffsd, "f", x), fx["ffn_out"], atol=1e-5)
    assert torch.equal(O.positional_encoding(40, fx["D"]), fx["pe"])


def test_compare_tokens_tracer():
    ref = torch.tensor([[1, 5, 6, 7], [1, 8, 9, 3]])
    logits = torch.zeros(2, 3, 10)
    logits[1, 1, 9] = 1.0
    logits[1, 1, 4] = 0.995      # near tie at the step that produced token index 2 of row 1
    got = ref.clone()
    got[1, 2] = 4
    r = O.compare_tokens(ref, logits, got)
    assert r["identical"] == 1 and len(r["near_tie"]) == 1 and not r["hard"]
    logits[1, 1, 4] = 0.5
    r = O.compare_tokens(ref, logits, got)
    assert len(r["hard"]) == 1


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules", "Transformer")), reason="reference tree not present")
def test_oracle_against_live_reference():
    sys.path.insert(0, REF)
    try:
        from modules.Transformer.model import Transformer as Ref
    finally:
        sys.path.remove(REF)
    cfg = O.CONFIGS["T0"]
    torch.manual_seed(0)
    ref = Ref(**cfg.ctor_kwargs()).eval()
    with torch.no_grad():
        for p in ref.parameters():
            O.bf16_representable_(p)
        sd = ref.state_dict()
        # our mirror modules draw the same RNG stream and expose the same keys
        ours = build_model(cfg)
        so = ours.state_dict()
        assert list(so.keys()) == list(sd.keys())
        assert all(torch.equal(so[k], sd[k]) for k in sd)
        ours.load_state_dict(sd, strict=True)
        spec = O.structured_spectrum(2, cfg.frames, cfg.input_dim, seed=7)
        text, mask = O.teacher_text(cfg, 2, seed=8)
        assert torch.allclose(O.transformer_forward(sd, spec, text, mask), ref(spec, text, mask), atol=2e-4)
        bos = torch.full((2, 1), cfg.bos_token_id, dtype=torch.int32)
        tok_ref, probs_ref = ref.evaluate(spec, bos)
        tok_o, probs_o, _ = O.evaluate_reference_style(sd, spec, cfg)
        assert torch.equal(tok_ref, tok_o) and len(probs_ref) == len(probs_o)
        assert all(torch.allclose(a, b, atol=2e-4) for a, b in zip(probs_ref, probs_o))


def test_beam_search_definition():
    """The oracle's beam search (the semantics the CUDA path is held to; the reference has none, README.md:30):
    beam 1 is the greedy search of model.py:125-151 padded after EOS, a wider beam never scores worse, hypotheses
    come out best first, and every reported score is the sum of the log-probabilities of its tokens."""
    cfg = O.CONFIGS["T0"]
    m = build_model(cfg)
    sd = cpu_state(m)
    spec = O.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=1)
    enc = O.encode(sd, spec)
    tg, lg = O.greedy_kv_cached(sd, enc, cfg)
    t1, s1 = O.beam_search_kv_cached(sd, enc, cfg, 1)
    for b in range(cfg.batch):
        ref = tg[b].clone()
        eos = (ref[1:] == cfg.eos_token_id).nonzero()
        if eos.numel():
            ref[int(eos[0]) + 2:] = cfg.pad_token_id
        assert torch.equal(t1[b, 0], ref)
        n = int(eos[0]) + 1 if eos.numel() else cfg.decoder_seq_len        # scored steps
        lp = torch.log_softmax(lg[b], -1)
        want = sum(float(lp[t, tg[b, t + 1]]) for t in range(n))
        assert abs(float(s1[b, 0]) - want) < 1e-3
    t4, s4 = O.beam_search_kv_cached(sd, enc, cfg, 4)
    assert t4.shape == (cfg.batch, 4, cfg.decoder_seq_len + 1)
    assert (s4[:, 0] >= s1[:, 0] - 1e-5).all() and (s4[:, :-1] >= s4[:, 1:]).all()
    assert len({tuple(r.tolist()) for r in t4[0]}) == 4                   # distinct hypotheses
