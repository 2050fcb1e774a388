"""Audio front-end: the oracle against torchaudio's Spectrogram (the object the reference's dataset
builds) and a literal DFT, and the GPU kernel against the oracle."""
import os

import pytest
import torch

from oracle import spectrogram as S

HERE = os.path.dirname(os.path.abspath(__file__))
FX = torch.load(os.path.join(HERE, "golden", "spectrogram.pt"), map_location="cpu", weights_only=False)


def close(got, ref, rtol=2e-4):
    scale = float(ref.abs().max())
    d = (got.double() - ref.double()).abs()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert float(d.max()) <= rtol * scale, f"max|d| {float(d.max()):.3e} vs scale {scale:.3e}"
    return float(d.max()) / scale


def test_oracle_matches_golden_and_dft():
    audio = S.synthetic_audio(FX["batch"], FX["n_samples"], FX["seed"])
    assert (FX["win_length"], FX["hop_length"]) == (1024, 512)          # torchaudio defaults the reference relies on
    close(S.spectrogram_ref(audio), FX["spec"])
    close(S.spectrogram_dft(audio[:1, :3000]).float(), S.spectrogram_ref(audio[:1, :3000]))
    padded = S.spectrogram_ref(audio, frames_out=20)
    assert padded.shape[-1] == 20 and not padded[..., 16:].any() and torch.equal(padded[..., :16], S.spectrogram_ref(audio))


def test_oracle_matches_live_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    tr = torchaudio.transforms.Spectrogram(n_fft=1024, center=False)
    audio = S.synthetic_audio(2, 5000, seed=9)
    close(S.spectrogram_ref(audio), tr(audio.unsqueeze(1)))


@pytest.mark.gpu
def test_gpu_spectrogram_golden_and_sizes():
    from asr_transformer_b200.frontend import Spectrogram
    sp = Spectrogram(n_fft=1024)
    audio = S.synthetic_audio(FX["batch"], FX["n_samples"], FX["seed"])
    got = sp(audio.cuda())
    assert got.shape == FX["spec"].shape and got.is_cuda
    close(got.cpu(), FX["spec"])
    # dataset-native shape: 10 s at 16 kHz -> 311 frames (the C0 config), zero padded to a fixed frame count
    audio = S.synthetic_audio(4, 160000, seed=6)
    ref = S.spectrogram_ref(audio, frames_out=320)
    got = sp(audio.cuda().unsqueeze(1), frames_out=320)
    assert got.shape == (4, 1, 513, 320) and sp.num_frames(160000) == 311
    close(got.cpu(), ref)
    assert not got[..., 311:].any()
    # ragged / edge cases: shorter than one window -> no frames; other FFT sizes
    assert sp(torch.zeros(2, 100, device="cuda")).shape == (2, 1, 513, 0)
    assert not sp(torch.zeros(2, 100, device="cuda"), frames_out=3).any()
    for n_fft in (256, 2048):
        a = S.synthetic_audio(2, 7000, seed=n_fft)
        close(Spectrogram(n_fft=n_fft)(a.cuda()).cpu(), S.spectrogram_ref(a, n_fft))
    assert sp(torch.zeros(0, 5000, device="cuda")).shape[0] == 0


@pytest.mark.gpu
def test_gpu_audio_to_tokens_no_host_round_trip():
    """audio -> spectrogram -> conv front-end -> encoder -> greedy decode, all on the device (C0 shape: 513 bins)."""
    from asr_transformer_b200.frontend import Spectrogram
    from tests.util import build_model, cpu_state
    from oracle import speech_transformer as O
    import dataclasses
    cfg = dataclasses.replace(O.CONFIGS["C0"], encoder_num_layers=1, decoder_num_layers=1, decoder_seq_len=8, batch=2)
    m = build_model(cfg, "cuda")
    audio = S.synthetic_audio(2, 160000, seed=7)
    spec_ref = torch.log1p(S.spectrogram_ref(audio))          # compress the dynamic range for the random-init model
    spec = torch.log1p(Spectrogram()(audio.cuda()))
    assert spec.shape == (2, 1, 513, 311)
    enc_ref = O.encode(cpu_state(m), O.bf16_representable_(spec_ref.clone()))
    enc = m.encode(O.bf16_representable_(spec.clone()))
    d = (enc.cpu() - enc_ref).abs()
    assert float(d.max()) < 6e-2 and float(d.mean()) < 6e-3
    tokens, _ = m.greedy_decode(spec)
    assert tokens.shape == (2, 9)


@pytest.mark.gpu
def test_gpu_transcribe_audio_to_text():
    """Transformer.transcribe = Spectrogram -> serving loop (stop at EOS) -> Detokenizer, batch by batch."""
    import dataclasses
    import json
    from asr_transformer_b200 import Detokenizer, Spectrogram
    from tests.util import build_model
    from oracle import speech_transformer as O
    cfg = dataclasses.replace(O.CONFIGS["C0"], encoder_num_layers=1, decoder_num_layers=1, decoder_seq_len=12, batch=2)
    m = build_model(cfg, "cuda")
    fx = json.load(open(os.path.join(HERE, "golden", "text_detok.json"), encoding="utf-8"))
    detok = Detokenizer(fx["vocab"], fx["special_ids"], fx["suffix"]) if "vocab" in fx else None
    if detok is None:
        pytest.skip("detokeniser fixture without an embedded vocabulary")
    sp = Spectrogram()
    audios = [S.synthetic_audio(2, 160000, seed=7), S.synthetic_audio(2, 160000, seed=8).cuda()]
    texts = list(m.transcribe(audios, sp, detok))
    assert len(texts) == 2 and all(len(t) == 2 and all(isinstance(x, str) for x in t) for t in texts)
    for a, got in zip(audios, texts):
        tok, n = m.greedy_decode(sp(a.cuda()), stop_at_eos=True)
        assert got == detok.decode_batch(tok.cpu(), n.cpu())
