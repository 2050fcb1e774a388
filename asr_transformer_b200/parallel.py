"""Multi-GPU host logic: one process per GPU, utterances sharded by batch, no collective on the compute path.

Utterances never interact (no cross-utterance op exists in reference model.py / layers.py), so the path shards
naturally: rank r decodes a contiguous slice of the batch with its own replica of the weights and its own KV
cache, and the only exchange is one final gather of the (tiny) token matrices.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n utterances: the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def balanced_assignment(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Mixed-length batches (BASELINE config 5): sort by length, deal round-robin so every rank gets a similar
    mix; returns the utterance indices per rank (each list sorted by decreasing length)."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    return [order[r::world] for r in range(world)]


def bucket_by_length(lengths: Sequence[int], batch_size: int) -> List[List[int]]:
    """Length-bucketed batching: utterance indices grouped into batches of ``batch_size``
    neighbours in length (longest first), so that a batch is padded to its own longest utterance instead of the
    longest of the corpus.  Every index appears exactly once; the last batch may be short."""
    if batch_size <= 0:
        raise ValueError("batch_size must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[i:i + batch_size] for i in range(0, len(order), batch_size)]


def gather_tokens(tokens: torch.Tensor, n_tokens: torch.Tensor, counts: Optional[Sequence[int]] = None,
                  group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Final transcript gather: every rank contributes (b_r, L+1) int32 tokens and (b_r,) lengths and receives the
    concatenation in rank order.  Ranks may hold different b_r (pass ``counts``, the per-rank sizes)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tokens, n_tokens
    world = dist.get_world_size(group)
    if counts is None:
        counts = [tokens.shape[0]] * world
    width = tokens.shape[1]
    bmax = max(counts) if len(counts) else 0
    packed = torch.zeros(bmax, width + 1, dtype=torch.int32, device=tokens.device)
    packed[:tokens.shape[0], :width] = tokens.to(torch.int32)
    packed[:tokens.shape[0], width] = n_tokens.to(torch.int32)
    out = torch.empty(world * bmax, width + 1, dtype=torch.int32, device=tokens.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    rows = [out[r * bmax:r * bmax + counts[r]] for r in range(world)]
    allrows = torch.cat(rows, 0)
    return allrows[:, :width].contiguous(), allrows[:, width].contiguous()


def decode_sharded(model, spectrum_cpu: torch.Tensor, device: torch.device, lengths: Optional[torch.Tensor] = None,
                   group=None, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
    """Greedy ASR of a host batch across all ranks of ``group``: each rank uploads and decodes its slice, then the
    transcripts are gathered.  Every rank returns the full (B, L+1) token matrix."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = spectrum_cpu.shape[0]
    lo, hi = shard_range(B, rank, world)
    counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    local = spectrum_cpu[lo:hi].to(device, non_blocking=True)
    loc_len = None if lengths is None else lengths[lo:hi].to(device)
    tokens, n_tok = model.greedy_decode(local, lengths=loc_len, **kw)
    return gather_tokens(tokens, n_tok, counts, group)
