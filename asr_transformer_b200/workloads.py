"""Named workloads and synthetic inputs of the hot path: the reference has no default
hyper-parameters (model.py:155-166), so the BASELINE configurations are pinned here; bench.py, the profiling tools, the
tests and the CPU oracle all take them from this one place.  Pure data + torch RNG on the host: no kernels, no oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Optional

import torch

Tensor = torch.Tensor


@dataclass(frozen=True)
class Config:
    name: str
    vocab_size: int = 250          # tokenizer.json:125-375
    input_dim: int = 80            # F (mel bins)
    embedding_dim: int = 256       # D
    decoder_seq_len: int = 128     # L
    encoder_seq_len: int = 249     # T' (must be >= conv_len(conv_len(T)), layers.py:73)
    encoder_num_layers: int = 6
    decoder_num_layers: int = 6
    num_heads: int = 4
    ff_dim: int = 1024
    pad_token_id: int = 4          # model.py:165
    eos_token_id: int = 2          # model.py:166
    bos_token_id: int = 1          # tokenizer.json:127
    frames: int = 1000             # T (10 ms hop)
    batch: int = 8

    def ctor_kwargs(self) -> dict:
        """kwargs of ``Transformer.__init__`` (model.py:155-166)."""
        d = asdict(self)
        for k in ("name", "bos_token_id", "frames", "batch"):
            d.pop(k)
        d["dropout"] = 0.1
        return d


def conv_len(n: int) -> int:
    """Output length of one 3-wide, stride-2, unpadded conv (model.py:168-172)."""
    return (n - 3) // 2 + 1


def subsampled_len(n: int) -> int:
    return conv_len(conv_len(n))


CONFIGS: Dict[str, Config] = {
    # tiny config for fast fixtures / smoke (not a BASELINE config)
    "T0": Config("T0", embedding_dim=128, num_heads=2, ff_dim=256, encoder_num_layers=2,
                 decoder_num_layers=2, decoder_seq_len=16, frames=200,
                 encoder_seq_len=subsampled_len(200), batch=3),
    "C1": Config("C1", batch=8),
    "C2": Config("C2", batch=64),
    "C3": Config("C3", encoder_num_layers=12, batch=256),
    "C4": Config("C4", encoder_num_layers=12, frames=3000, encoder_seq_len=subsampled_len(3000),
                 decoder_seq_len=384, batch=64),
    "C5": Config("C5", embedding_dim=512, num_heads=8, ff_dim=2048, encoder_num_layers=12, batch=64),
    "C0": Config("C0", input_dim=513, frames=311, encoder_seq_len=subsampled_len(311), batch=8),
}


# --------------------------------------------------------------------------
# synthetic inputs / weights
# --------------------------------------------------------------------------
def bf16_representable_(t: Tensor) -> Tensor:
    """Round in place to the nearest bf16-representable fp32 value."""
    t.copy_(t.to(torch.bfloat16).to(torch.float32))
    return t


def structured_spectrum(batch: int, frames: int, input_dim: int = 80, seed: int = 1,
                        lengths: Optional[Tensor] = None) -> Tensor:
    """Structured 'mel-ish' synthetic input (B,1,F,T), bf16-representable fp32.

    i.i.d. randn inputs make every utterance decode to the same tokens at random
    init; this generator gives distinct rows.
    """
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(frames, dtype=torch.float32)[None, None, :]
    f = torch.arange(input_dim, dtype=torch.float32)[None, :, None]
    ph = torch.rand(batch, 1, 1, generator=g) * (2 * math.pi)
    fr = 0.005 + 0.05 * torch.rand(batch, 1, 1, generator=g)
    noise = torch.randn(batch, input_dim, frames, generator=g)
    x = -4.0 + 3.0 * torch.sin(2 * math.pi * fr * t + ph + 0.1 * f) \
        + 2.0 * torch.cos(0.2 * f * (1.0 + ph)) + 0.5 * noise
    if lengths is not None:   # zero-pad the time axis as dataset.py:53-55 does
        keep = torch.arange(frames)[None, None, :] < lengths.view(-1, 1, 1)
        x = x * keep
    return bf16_representable_(x.unsqueeze(1).contiguous())




def build_model(cfg: Config, device="cpu"):
    """The drop-in Transformer with the reference's seed-0 default init (same parameter construction order, so the
    state_dict equals the reference's bit for bit), rounded to bf16-representable fp32, in eval mode."""
    from .model import Transformer
    torch.manual_seed(0)
    m = Transformer(**cfg.ctor_kwargs())
    with torch.no_grad():
        for p in m.parameters():
            bf16_representable_(p)
    m.eval()
    return m.to(device)


def cpu_state(m) -> Dict[str, Tensor]:
    """fp32 CPU copy of a module's state_dict (what the CPU oracle consumes)."""
    return {k: v.detach().cpu().float() if v.is_floating_point() else v.detach().cpu()
            for k, v in m.state_dict().items()}
