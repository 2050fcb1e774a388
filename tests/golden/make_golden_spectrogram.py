#!/usr/bin/env python
"""Golden vectors for the audio front-end: torchaudio.transforms.Spectrogram(n_fft=1024, center=False), the exact
object reference modules/dataset.py:34-35 constructs, applied to seeded synthetic audio.  Run in the build container
(torchaudio installed); writes tests/golden/spectrogram.pt (audio is regenerated from the seed, only outputs stored)."""
import os
import sys

import torch
import torchaudio

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import spectrogram as S  # noqa: E402

tr = torchaudio.transforms.Spectrogram(n_fft=1024, center=False)
audio = S.synthetic_audio(3, 9000, seed=5)
ref = tr(audio.unsqueeze(1))                                   # (B, 1, 513, 16)
ours = S.spectrogram_ref(audio)
assert ref.shape == ours.shape == (3, 1, 513, S.n_frames(9000)), (ref.shape, ours.shape)
assert torch.allclose(ref, ours, rtol=1e-4, atol=1e-4 * float(ref.max()))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "spectrogram.pt")
torch.save({"seed": 5, "batch": 3, "n_samples": 9000, "spec": ref.clone(),
            "win_length": tr.win_length, "hop_length": tr.hop_length}, out)
print("wrote", out, tuple(ref.shape), "max", float(ref.max()))
