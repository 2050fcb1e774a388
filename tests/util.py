"""Shared helpers for the parity tests."""
import hashlib
import os

import torch

from oracle import speech_transformer as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# stated tolerances: bf16 tensor-core operands, fp32 accumulate / epilogue / residual
TOL_MAX = 3e-2     # max |delta| for enc_out and logits
TOL_MEAN = 5e-3    # mean |delta|
TOL_FP32 = 1e-5    # fp32-only kernels (LayerNorm, PE, embedding)
TAU = 2e-2         # near-tie threshold on the fp32 reference top1-top2 margin


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def state_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


from asr_transformer_b200.workloads import build_model, cpu_state  # noqa: E402,F401


def assert_close(got, ref, tol_max=TOL_MAX, tol_mean=TOL_MEAN, what=""):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    d = (got - ref).abs()
    assert d.max().item() <= tol_max and d.mean().item() <= tol_mean, \
        f"{what}: max|d|={d.max().item():.3e} (tol {tol_max}), mean|d|={d.mean().item():.3e} (tol {tol_mean})"
    return d.max().item(), d.mean().item()
