"""Generate the golden fixtures by executing the UNMODIFIED reference (read-only, /root/reference).

Run in the build container only (the GPU box has no reference tree):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures - outputs of the
reference's own PyTorch modules on seeded weights and inputs - are what pins the oracle and the CUDA path.
Weights are NOT stored: they are regenerated from the seed by constructing the model under torch.manual_seed(0)
(our mirror modules draw the same RNG stream as the reference's, checked bit-exactly here and by a checksum in
the tests) and rounded to bf16-representable fp32.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("ASR_REF", "/root/reference"))

from modules.Transformer import layers as ref_layers          # noqa: E402  (reference, read-only)
from modules.Transformer.model import Transformer as RefTransformer   # noqa: E402

from oracle import speech_transformer as O                    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def state_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference(cfg: O.Config):
    torch.manual_seed(0)
    ref = RefTransformer(**cfg.ctor_kwargs())
    with torch.no_grad():
        for p in ref.parameters():
            O.bf16_representable_(p)
    ref.eval()
    return ref


def per_utterance_tokens(ref, spec, cfg):
    rows, last_probs = [], []
    for b in range(spec.shape[0]):
        bos = torch.full((1, 1), cfg.bos_token_id, dtype=torch.int32)
        tok, probs = ref.evaluate(spec[b:b + 1], bos)
        rows.append(tok)
        last_probs.append(probs[-1])      # (L-1, V): logits of positions 0..L-2 recomputed at the last step
    return torch.cat(rows, 0), torch.stack(last_probs, 0)


def model_fixture(name: str, with_forward: bool = True):
    cfg = O.CONFIGS[name]
    t0 = time.time()
    ref = build_reference(cfg)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    spec = O.structured_spectrum(cfg.batch, cfg.frames, cfg.input_dim, seed=1)
    fx = {"config": name, "state_checksum": state_checksum(sd)}
    with torch.no_grad():
        conv = ref.input_layer(spec)
        enc = ref.encoder(conv)
        fx["enc_out"] = enc.clone()
        if name == "T0":
            fx["conv_out"] = conv.clone()
        if with_forward:
            text, mask = O.teacher_text(cfg, cfg.batch, seed=2)
            fx["text"], fx["mask"] = text, mask
            fx["forward_logits"] = ref(spec, text, mask).clone()
        tokens, last_probs = per_utterance_tokens(ref, spec, cfg)
        fx["tokens"] = tokens                       # (B, L+1) int64, per-utterance reference greedy
        fx["last_probs"] = last_probs               # (B, L-1, V)
        # batch call: the reference's own return contract (last sample only + probs list)
        bos = torch.full((cfg.batch, 1), cfg.bos_token_id, dtype=torch.int32)
        tok_last, probs = ref.evaluate(spec, bos)
        fx["evaluate_tokens_last"] = tok_last
        fx["evaluate_probs_shapes"] = [tuple(p.shape) for p in probs]
        fx["evaluate_probs_sum"] = torch.tensor([float(p.double().sum()) for p in probs])
        # oracle pinned against the reference
        o_enc = O.encode(sd, spec)
        o_tok, o_logits = O.greedy_kv_cached(sd, o_enc, cfg)
        assert torch.allclose(o_enc, enc, atol=2e-5, rtol=1e-5), (o_enc - enc).abs().max()
        assert torch.equal(o_tok, tokens), "oracle KV-cached greedy differs from the reference"
        assert torch.allclose(o_logits[:, :-1], last_probs, atol=2e-4, rtol=1e-4), \
            (o_logits[:, :-1] - last_probs).abs().max()
        fx["step_logits"] = o_logits                # (B, L, V) fp32 (oracle; equals reference probs to 2e-4)
        if with_forward:
            o_fwd = O.decoder_forward(sd, text, mask, o_enc)
            assert torch.allclose(o_fwd, fx["forward_logits"], atol=2e-4, rtol=1e-4)
    fx["distinct_rows"] = len({tuple(r.tolist()) for r in tokens})
    torch.save(fx, os.path.join(OUT, f"model_{name}.pt"))
    print(f"{name}: distinct token rows {fx['distinct_rows']}/{cfg.batch}, {time.time() - t0:.1f}s, "
          f"checksum {fx['state_checksum'][:12]}")


def mha_fixture():
    """Operator-level oracle: reference layers.MHA with key-padding, causal and fully-masked rows (Q6, Q7)."""
    torch.manual_seed(3)
    D, H, B, Sq, Sk = 128, 2, 3, 37, 53
    mha = ref_layers.MHA(H, D, 0.1).eval()
    with torch.no_grad():
        for p in mha.parameters():
            O.bf16_representable_(p)
    g = torch.Generator().manual_seed(4)
    x = O.bf16_representable_(torch.randn(B, Sq, D, generator=g))
    src = O.bf16_representable_(torch.randn(B, Sk, D, generator=g))
    klen = torch.tensor([53, 20, 1])
    keypad = (torch.arange(Sk)[None, None, :] >= klen[:, None, None]).expand(B, Sq, Sk).clone()
    causal = torch.triu(torch.ones(Sq, Sq, dtype=torch.uint8), diagonal=1)
    full_rows = torch.zeros(B, Sq, Sq, dtype=torch.bool)
    full_rows[:, 5] = True            # fully masked query rows -> output == _out_linear.bias (Q7)
    full_rows[1, :, 10:] = True
    with torch.no_grad():
        fx = {
            "state": {k: v.clone() for k, v in mha.state_dict().items()},
            "D": D, "H": H, "x": x, "src": src, "keypad": keypad, "causal": causal, "full_rows": full_rows,
            "self_nomask": mha(x), "cross_nomask": mha(x, src), "cross_keypad": mha(x, src, keypad),
            "self_causal": mha(x, attention_mask=causal), "self_full_rows": mha(x, attention_mask=full_rows),
        }
        ff = ref_layers.FeedForward(D, 256, 0.1).eval()
        for p in ff.parameters():
            O.bf16_representable_(p)
        fx["ffn_state"] = {k: v.clone() for k, v in ff.state_dict().items()}
        fx["ffn_out"] = ff(x)
        pe = ref_layers.TrainablePositionalEncoding(40, D)
        fx["pe"] = pe.pe.clone()
        # oracle pinned against the reference operator
        sd = {"m." + k: v for k, v in fx["state"].items()}
        for key, args in (("self_nomask", (x, None, None)), ("cross_keypad", (x, src, keypad)),
                          ("self_causal", (x, None, causal)), ("self_full_rows", (x, None, full_rows))):
            assert torch.allclose(O.mha(sd, "m", *args), fx[key], atol=1e-5), key
        assert torch.equal(O.positional_encoding(40, D), fx["pe"])
    torch.save(fx, os.path.join(OUT, "ops_mha.pt"))
    print("ops_mha: ok")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    mha_fixture()
    model_fixture("T0")
    model_fixture("C1", with_forward=True)
