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


class PeerGather:
    """All-gather of the packed per-sequence results by peer-to-peer stores over NVLink / NVSwitch issued by ONE tiny
    kernel of this library (`specdec_peer_publish`, csrc/engine.cu) right behind the verify step on the verify stream:
    no NCCL kernel competes with the persistent row kernel for an SM slot, no collective is enqueued per step, and the
    result lands in every rank's gather buffer `latency of one NVLink store + flag` after the step ends.

    The buffers are symmetric memory (torch.distributed._symmetric_memory: peer-mapped allocations exchanged once at
    construction, the only collective this class ever runs).  A ring of `slots` buffers lets a rank run up to
    `slots - 1` steps ahead of the slowest reader; `gathered(step)` enqueues a one-warp wait for that step's arrival
    flags on the current stream and returns the [total, width] view.
    """

    def __init__(self, total: int, width: int, group=None, slots: int = 8, device=None, overlap: bool = False):
        """overlap=True: the publish kernel runs on a side stream of this object behind an event recorded on the caller's
        stream (as NCCL's own stream does), so it never sits on the verify stream's critical path; call
        `sync_reader()` before consuming a gathered view on the caller's stream."""
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self._L = L
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        self.total, self.width, self.slots = int(total), int(width), int(slots)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.lo, self.hi = shard_range(self.total, self.rank, self.world)
        self.buf = symm.empty((self.slots, self.total, self.width), dtype=torch.int32, device=dev)
        self.flags = symm.empty((self.slots, max(self.world, 4)), dtype=torch.int32, device=dev)
        self.buf.fill_(-1)
        self.flags.zero_()
        torch.cuda.synchronize(dev)
        hb = symm.rendezvous(self.buf, grp.group_name)
        hf = symm.rendezvous(self.flags, grp.group_name)
        self._peer_bufs = torch.tensor([int(p) for p in hb.buffer_ptrs], dtype=torch.int64, device=dev)
        self._peer_flags = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flags_ptr = self.flags.data_ptr()
        self.side = torch.cuda.Stream(device=dev) if overlap else None
        self._ev = [torch.cuda.Event() for _ in range(4)] if overlap else None
        self._keep = [None] * 4  # the published blocks stay referenced until the side stream has certainly read them
        self._handles = (hb, hf)
        dist.barrier(grp)
        torch.cuda.synchronize(dev)

    def _slot_seq(self, step: int):
        return step % self.slots, step // self.slots + 1

    def publish(self, packed_local: torch.Tensor, step: int, wait_step: int = -1):
        """packed_local: this rank's int32 [hi - lo, width] rows of step `step` (contiguous, CUDA).  wait_step >= 0: the
        same launch also waits (in stream order) for every rank's results of that EARLIER step and their [total, width]
        view is returned -- one kernel per step for a reader that lags the writers."""
        if packed_local.dtype != torch.int32 or not packed_local.is_contiguous() or packed_local.numel() != (self.hi - self.lo) * self.width:
            raise ValueError("PeerGather.publish: expected the rank's contiguous int32 [rows, width] block")
        slot, seq = self._slot_seq(step)
        stream = torch.cuda.current_stream().cuda_stream
        if self.side is not None:
            ev = self._ev[step & 3]
            ev.record()
            self.side.wait_event(ev)
            self._keep[step & 3] = packed_local
            stream = self.side.cuda_stream
        rc = self._L.lib().specdec_peer_publish(
            packed_local.data_ptr(), packed_local.numel(), self._peer_bufs.data_ptr(),
            (slot * self.total + self.lo) * self.width, self.world, self._peer_flags.data_ptr(),
            slot * self.flags.shape[1] + self.rank, seq,
            (self._flags_ptr + 4 * (wait_step % self.slots) * self.flags.shape[1]) if wait_step >= 0 else None,
            (wait_step // self.slots + 1) if wait_step >= 0 else 0, self.status.data_ptr(), stream)
        self._L.check(rc, "specdec_peer_publish")
        return self.buf[wait_step % self.slots] if wait_step >= 0 else None

    def sync_reader(self) -> None:
        """overlap mode: the caller's stream waits for everything published / awaited so far on the side stream."""
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)

    def gathered(self, step: int) -> torch.Tensor:
        """[total, width] results of every rank for `step`, valid in stream order after this call."""
        slot, seq = self._slot_seq(step)
        rc = self._L.lib().specdec_peer_wait(self.flags[slot].data_ptr(), self.world, seq, self.status.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        self._L.check(rc, "specdec_peer_wait")
        return self.buf[slot]
