"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's audio front-end (never imported by the product).

Reference: modules/dataset.py:34-35 builds ``torchaudio.transforms.Spectrogram(n_fft=1024, center=False)`` and :51-55
applies it and zero-pads the frame axis.  torchaudio's defaults (win_length = n_fft, hop_length = win_length // 2,
periodic Hann window, power = 2, one-sided, not normalised) make that transform

    X[k, t] = | sum_n  hann[n] * x[t * hop + n] * exp(-2 pi i k n / n_fft) |^2 ,   k = 0 .. n_fft/2,
    t = 0 .. floor((N - n_fft) / hop)

``spectrogram_ref`` restates it with torch.stft (what torchaudio calls); ``spectrogram_dft`` is the literal float64
sum, independent of any FFT library.  Pinned by tests/test_frontend.py against torchaudio itself (installed in the
build container) and by the committed golden fixture tests/golden/spectrogram.pt.
"""
from __future__ import annotations

import math

import torch


def n_frames(n_samples: int, n_fft: int = 1024, hop: int | None = None) -> int:
    hop = hop or n_fft // 2
    return (n_samples - n_fft) // hop + 1 if n_samples >= n_fft else 0


def spectrogram_ref(audio: torch.Tensor, n_fft: int = 1024, hop: int | None = None, frames_out: int | None = None):
    """audio (B, N) fp32 -> (B, 1, n_fft/2+1, T) fp32; T = frames_out (zero padded, dataset.py:53-55) or the frame count."""
    hop = hop or n_fft // 2
    win = torch.hann_window(n_fft, periodic=True, dtype=audio.dtype)
    X = torch.stft(audio, n_fft, hop_length=hop, win_length=n_fft, window=win, center=False, normalized=False,
                   onesided=True, return_complex=True)
    P = (X.real ** 2 + X.imag ** 2).unsqueeze(1)
    if frames_out is not None and frames_out > P.shape[-1]:
        P = torch.cat([P, P.new_zeros(*P.shape[:-1], frames_out - P.shape[-1])], -1)
    return P


def spectrogram_dft(audio: torch.Tensor, n_fft: int = 1024, hop: int | None = None):
    """Literal float64 DFT sum (small inputs only)."""
    hop = hop or n_fft // 2
    x = audio.double()
    T = n_frames(x.shape[-1], n_fft, hop)
    n = torch.arange(n_fft, dtype=torch.float64)
    win = 0.5 - 0.5 * torch.cos(2 * math.pi * n / n_fft)
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)
    ang = -2 * math.pi * k[:, None] * n[None, :] / n_fft
    C, S = torch.cos(ang), torch.sin(ang)
    frames = torch.stack([x[:, t * hop:t * hop + n_fft] * win for t in range(T)], 1)     # (B, T, n_fft)
    re = torch.einsum("btn,kn->bkt", frames, C)
    im = torch.einsum("btn,kn->bkt", frames, S)
    return (re ** 2 + im ** 2).unsqueeze(1)


def synthetic_audio(batch: int, n_samples: int, seed: int = 5) -> torch.Tensor:
    """Chirps + noise in [-1, 1], different per utterance."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n_samples, dtype=torch.float32) / 16000.0
    f0 = 100 + 400 * torch.rand(batch, 1, generator=g)
    sweep = 2000 * torch.rand(batch, 1, generator=g)
    x = 0.6 * torch.sin(2 * math.pi * (f0 * t + 0.5 * sweep * t * t)) + 0.05 * torch.randn(batch, n_samples, generator=g)
    return x.clamp(-1, 1)
