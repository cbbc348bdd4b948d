"""Request-sharded multi-GPU plumbing: one process per GPU, sequences sharded
contiguously across ranks, no data-path collective; after each verify step ONE all-gather of the
packed int32 [B_local, gamma+2] result ({n, accepted tokens, next token}) over NCCL/NVLink gives every
rank all sequences' outcomes.  Philox uniforms are keyed by the GLOBAL sequence id, so results do
not depend on the number of ranks.  The same code runs on gloo/CPU tensors for the host-logic tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` sequences owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(n_accepted: torch.Tensor, draft_tokens: torch.Tensor, next_token: torch.Tensor) -> torch.Tensor:
    """Host-side mirror of the kernel's `packed` output (used by the gloo tests)."""
    B, g = draft_tokens.shape
    out = torch.full((B, g + 2), -1, dtype=torch.int32, device=draft_tokens.device)
    out[:, 0] = n_accepted.to(torch.int32)
    ar = torch.arange(g, device=draft_tokens.device).unsqueeze(0)
    keep = ar < n_accepted.long().unsqueeze(1)
    out[:, 1:g + 1] = torch.where(keep, draft_tokens.to(torch.int32), out[:, 1:g + 1])
    out.scatter_(1, (n_accepted.long() + 1).unsqueeze(1), next_token.to(torch.int32).unsqueeze(1))
    return out


def all_gather_packed(packed_local: torch.Tensor, total: int, group=None, async_op: bool = False):
    """All-gather the packed per-sequence results of every rank -> [total, gamma+2] (global order).
    async_op=True (even shards only) returns (out, work): the collective runs on NCCL's own stream so the
    next verify step overlaps it; call work.wait() before reading `out`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (packed_local, None) if async_op else packed_local
    world = dist.get_world_size(group)
    width = packed_local.shape[1]
    base, rem = divmod(total, world)
    if rem == 0:  # even shards: one all_gather_into_tensor (single NCCL all-gather)
        out = torch.empty((total, width), dtype=packed_local.dtype, device=packed_local.device)
        work = dist.all_gather_into_tensor(out, packed_local.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    assert not async_op, "async all-gather needs evenly divisible shards"
    mx = base + 1
    pad = torch.full((mx, width), -1, dtype=packed_local.dtype, device=packed_local.device)
    pad[:packed_local.shape[0]] = packed_local
    out = torch.empty((world * mx, width), dtype=packed_local.dtype, device=packed_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        parts.append(out[r * mx:r * mx + (hi - lo)])
    return torch.cat(parts, 0)


def unpack_results(packed: torch.Tensor):
    """-> (n_accepted [B], tokens [B, gamma+1] with -1 padding; tokens[b, n_b] is the next token)"""
    return packed[:, 0], packed[:, 1:]
