"""Audio front-end on the GPU: raw audio -> power spectrogram in the layout the conv front-end reads.

Mirrors ``torchaudio.transforms.Spectrogram(n_fft, center=False)`` as constructed by the reference's dataset
(modules/dataset.py:34-35): win_length = n_fft, hop_length = n_fft // 2, periodic Hann window, power 2, one-sided.
``frames_out`` additionally zero-pads the frame axis on the device like dataset.py:53-55 does on the host.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import lib as _l


class Spectrogram(nn.Module):
    def __init__(self, n_fft: int = 1024, win_length: Optional[int] = None, hop_length: Optional[int] = None,
                 power: float = 2.0, center: bool = False):
        super().__init__()
        if win_length not in (None, n_fft) or power != 2.0 or center:
            raise ValueError("asr_b200 Spectrogram implements the reference's configuration only: win_length = n_fft, "
                             "power = 2, center = False (modules/dataset.py:34-35)")
        self.n_fft = n_fft
        self.win_length = n_fft
        self.hop_length = hop_length if hop_length is not None else n_fft // 2

    def num_frames(self, n_samples: int) -> int:
        """dataset.py:41-42 (_calculate_spectrum_len)."""
        return (n_samples - self.win_length) // self.hop_length + 1 if n_samples >= self.win_length else 0

    def forward(self, audio: torch.Tensor, frames_out: Optional[int] = None) -> torch.Tensor:
        """audio (B, N) or (B, 1, N) fp32 CUDA -> (B, 1, n_fft // 2 + 1, T) fp32, T = frames_out or the frame count."""
        if audio.dim() == 3 and audio.shape[1] == 1:
            audio = audio[:, 0]
        if audio.dim() != 2:
            raise RuntimeError(f"expected audio (B, N) or (B, 1, N), got {tuple(audio.shape)}")
        audio = audio.to(torch.float32).contiguous()
        B, N = audio.shape
        T = int(frames_out) if frames_out is not None else self.num_frames(N)
        out = torch.empty(B, 1, self.n_fft // 2 + 1, max(T, 0), dtype=torch.float32, device=audio.device)
        if B and T > 0:
            _l.check(_l.load().asr_spectrogram(_l.ptr(audio), B, N, self.n_fft, self.hop_length, T, _l.ptr(out),
                                               _l.stream()), "asr_spectrogram")
        return out
