"""Batch sharding of the tokenizer across the GPUs of one box.

Clouds are independent in eval mode, so the path shards by batch with no collective:
rank r tokenizes clouds [lo, hi) of the global batch.  The only communication is the optional final
all-gather of token shards (BASELINE.json configs[4]) - plain NCCL over NVLink/NVSwitch; at <= 50 MB per
rank it is launch-latency sized, not worth a custom kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_slice(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `batch` clouds for `rank` (first `batch % world` ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_slice(int(x.shape[0]), rank, world)
    return x[lo:hi]


def gather_tokens(tokens: torch.Tensor, batch: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gather per-rank token shards (B_r, G, E) into the global (B, G, E) tensor on every rank.
    Ragged shards (batch % world != 0) are padded to the largest shard and trimmed after the gather."""
    world = dist.get_world_size(group)
    if world == 1:
        return tokens
    n_local = torch.tensor([tokens.shape[0]], dtype=torch.int64, device=tokens.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c) for c in counts]
    mx = max(counts)
    if tokens.shape[0] < mx:
        pad = tokens.new_zeros((mx - tokens.shape[0],) + tuple(tokens.shape[1:]))
        tokens = torch.cat([tokens, pad], 0)
    out = tokens.new_empty((world * mx,) + tuple(tokens.shape[1:]))
    dist.all_gather_into_tensor(out, tokens.contiguous(), group=group)
    if all(c == mx for c in counts):
        return out
    parts = [out[r * mx:r * mx + c] for r, c in enumerate(counts)]
    return torch.cat(parts, 0)
