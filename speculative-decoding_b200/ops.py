"""torch.library custom ops over the C ABI (host side stays Python/PyTorch; PyTorch is only
device memory + streams here).  Every op requires CUDA tensors and the built library; there is
no eager / CPU implementation behind them."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float16: L.F16}


def _dtype_code(t: Tensor) -> int:
    if t.dtype not in _DT:
        raise TypeError(f"specdec: unsupported logits dtype {t.dtype} (float32 / bfloat16 / float16)")
    return _DT[t.dtype]


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("specdec ops run on CUDA tensors only (no CPU fallback)")


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rows_view(logits: Tensor) -> Tuple[Tensor, int, int, int]:
    """[..., V] -> uniform row stride view (rows, V, stride)."""
    V = logits.shape[-1]
    x = logits if logits.stride(-1) == 1 else logits.contiguous()
    x = x.reshape(-1, V)
    if x.stride(-1) != 1:
        x = x.contiguous()
    return x, x.shape[0], V, (x.stride(0) if x.shape[0] > 1 else V)


_WS_CACHE = {}


_WS_BYTES = {}


def _workspace(dev, B_or_bytes: int, gamma: int = -1, V: int = 0, lib=None) -> Tensor:
    """Per-device-and-stream scratch, grown on demand and reused across calls (all use is stream-ordered).
    Called with (dev, nbytes) or, for verify, with (dev, B, gamma, V, lib): the size query is cached per shape."""
    if gamma >= 0:
        k = (B_or_bytes, gamma, V)
        nbytes = _WS_BYTES.get(k)
        if nbytes is None:
            nbytes = _WS_BYTES[k] = lib.specdec_verify_workspace_bytes(B_or_bytes, gamma, V)
    else:
        nbytes = B_or_bytes
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = ws
    return ws


def _verify_impl(target_logits: Tensor, draft_logits: Optional[Tensor], draft_tokens: Tensor,
                 u_accept: Optional[Tensor], u_sample: Optional[Tensor], seed: int, offset: int, seq_id0: int,
                 temperature: float, top_k: int, top_p: float, sample_mode: int, flags: int,
                 stop_tokens: Optional[Tensor]) -> List[Tensor]:
    _need_cuda(target_logits, draft_logits, draft_tokens, u_accept, u_sample, stop_tokens)
    if target_logits.dim() != 3:
        raise ValueError("target_logits must be [B, gamma(+1), V]")
    B, gT, V = target_logits.shape
    gamma = gT if (flags & L.NO_BONUS) else gT - 1
    dev = target_logits.device
    if target_logits.stride(-1) != 1:
        target_logits = target_logits.contiguous()
    dt = _dtype_code(target_logits)
    if draft_logits is not None:
        if draft_logits.dtype != target_logits.dtype:
            draft_logits = draft_logits.to(target_logits.dtype)
        if draft_logits.shape != (B, gamma, V):
            raise ValueError(f"draft_logits must be [B={B}, gamma={gamma}, V={V}], got {tuple(draft_logits.shape)}")
        if draft_logits.stride(-1) != 1:
            draft_logits = draft_logits.contiguous()
    if draft_tokens.dtype != torch.int64 or draft_tokens.device != dev or not draft_tokens.is_contiguous() \
            or draft_tokens.numel() != B * gamma:
        draft_tokens = draft_tokens.to(device=dev, dtype=torch.int64).reshape(B, gamma).contiguous()
    if u_accept is not None and (u_accept.dtype != torch.float32 or u_accept.device != dev or not u_accept.is_contiguous()):
        u_accept = u_accept.to(device=dev, dtype=torch.float32).reshape(B, gamma).contiguous()
    if u_sample is not None and (u_sample.dtype != torch.float32 or u_sample.device != dev or not u_sample.is_contiguous()):
        u_sample = u_sample.to(device=dev, dtype=torch.float32).reshape(B).contiguous()
    n_stop = 0
    if stop_tokens is not None:
        if stop_tokens.dtype != torch.int64 or stop_tokens.device != dev or not stop_tokens.is_contiguous():
            stop_tokens = stop_tokens.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        n_stop = stop_tokens.numel()
    # one allocation for all outputs: [n_acc i32 | first_stop i32 | next_prob f32 | p_tok f32 | q_tok f32 |
    #                                  packed i32 | next_token i64 (8-aligned) | mask u8]
    # Only addresses are computed here; the tensor views are created lazily by VerifyResult (a dozen view objects
    # cost more host time than the whole enqueue at small batch).
    g = gamma
    n32 = _n32(B, g)
    buf = torch.empty(n32 * 4 + B * 8 + B * g, dtype=torch.uint8, device=dev)
    base = buf.data_ptr()
    o_pt = base + 12 * B
    o_qt = o_pt + 4 * B * g
    o_pk = o_qt + 4 * B * g
    o_nx = base + n32 * 4
    o_mk = o_nx + 8 * B
    lib = L.lib()
    ws = _workspace(dev, B, gamma, V, lib)
    sd = draft_logits.stride() if draft_logits is not None else (0, 0, 1)
    if torch.cuda.current_device() != dev.index:
        torch.cuda.set_device(dev)
    rc = lib.specdec_verify(
        target_logits.data_ptr(), _ptr(draft_logits), dt, draft_tokens.data_ptr(), _ptr(u_accept), _ptr(u_sample),
        seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, seq_id0, B, gamma, V,
        target_logits.stride(0), target_logits.stride(1), sd[0], sd[1],
        float(temperature), int(top_k), float(top_p), int(sample_mode), int(flags),
        _ptr(stop_tokens), n_stop, base, o_nx, o_mk if g else None,
        o_pt if g else None, o_qt if g else None, base + 4 * B,
        base + 8 * B, o_pk, ws.data_ptr(), ws.numel(), _stream())
    L.check(rc, "specdec_verify")
    return buf, B, g


def _n32(B: int, g: int) -> int:
    n32 = 3 * B + 2 * B * g + B * (g + 2)
    return n32 + (n32 & 1)


def _verify_views(buf: Tensor, B: int, g: int) -> List[Tensor]:
    n32 = _n32(B, g)
    i32 = buf[:n32 * 4].view(torch.int32)
    f32 = buf[:n32 * 4].view(torch.float32)
    o = 3 * B
    p_tok = f32[o:o + B * g].view(B, g); o += B * g
    q_tok = f32[o:o + B * g].view(B, g); o += B * g
    packed = i32[o:o + B * (g + 2)].view(B, g + 2)
    nxt = buf[n32 * 4:n32 * 4 + B * 8].view(torch.int64)
    mask = buf[n32 * 4 + B * 8:].view(B, g)
    return [i32[0:B], nxt, mask, p_tok, q_tok, i32[B:2 * B], f32[2 * B:3 * B], packed]


@torch.library.custom_op("specdec::verify", mutates_args=())
def verify_op(target_logits: Tensor, draft_logits: Optional[Tensor], draft_tokens: Tensor,
              u_accept: Optional[Tensor], u_sample: Optional[Tensor], seed: int, offset: int, seq_id0: int,
              temperature: float, top_k: int, top_p: float, sample_mode: int, flags: int,
              stop_tokens: Optional[Tensor]) -> List[Tensor]:
    # custom ops may not return views of one buffer: clone the (tiny) outputs
    return [t.clone() for t in _verify_views(*_verify_impl(target_logits, draft_logits, draft_tokens, u_accept, u_sample, seed,
                                                           offset, seq_id0, temperature, top_k, top_p, sample_mode, flags,
                                                           stop_tokens))]


@verify_op.register_fake
def _(target_logits, draft_logits, draft_tokens, u_accept, u_sample, seed, offset, seq_id0, temperature, top_k,
      top_p, sample_mode, flags, stop_tokens):
    B, gT, V = target_logits.shape
    gamma = gT if (flags & L.NO_BONUS) else gT - 1
    e = target_logits.new_empty
    return [e(B, dtype=torch.int32), e(B, dtype=torch.int64), e((B, gamma), dtype=torch.uint8),
            e((B, gamma), dtype=torch.float32), e((B, gamma), dtype=torch.float32), e(B, dtype=torch.int32),
            e(B, dtype=torch.float32), e((B, gamma + 2), dtype=torch.int32)]


@torch.library.custom_op("specdec::process_probs", mutates_args=())
def process_probs_op(logits: Tensor, temperature: float, top_k: int, top_p: float) -> List[Tensor]:
    """-> [probs fp32 [..., V], row_stats [rows, 8]]  (LogitsProcessor.__call__, utils/logits_processor.py:13-15)"""
    _need_cuda(logits)
    x, rows, V, stride = _rows_view(logits)
    dev = logits.device
    probs = torch.empty((rows, V), dtype=torch.float32, device=dev)
    stats = torch.empty((rows, 8), dtype=torch.float32, device=dev)
    lib = L.lib()
    ws_bytes = lib.specdec_workspace_bytes(rows)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.specdec_process_probs(_ptr(x), _dtype_code(x), rows, V, stride, float(temperature), int(top_k),
                                       float(top_p), _ptr(probs), _ptr(stats), _ptr(ws), ws_bytes, _stream())
    L.check(rc, "specdec_process_probs")
    return [probs.reshape(logits.shape), stats]


@process_probs_op.register_fake
def _(logits, temperature, top_k, top_p):
    rows = logits.numel() // logits.shape[-1]
    return [logits.new_empty(logits.shape, dtype=torch.float32), logits.new_empty((rows, 8), dtype=torch.float32)]


@torch.library.custom_op("specdec::sample_rows", mutates_args=())
def sample_rows_op(logits: Tensor, u: Optional[Tensor], seed: int, offset: int, seq_id0: int, lane_id: int,
                   temperature: float, top_k: int, top_p: float, sample_mode: int) -> List[Tensor]:
    """processor + sample() fused: -> [tok int64 [rows], ptok fp32 [rows]]"""
    _need_cuda(logits, u)
    x, rows, V, stride = _rows_view(logits)
    dev = logits.device
    if u is not None:
        u = u.to(device=dev, dtype=torch.float32).reshape(rows).contiguous()
    tok = torch.empty(rows, dtype=torch.int64, device=dev)
    ptok = torch.empty(rows, dtype=torch.float32, device=dev)
    lib = L.lib()
    ws_bytes = lib.specdec_sample_rows_workspace_bytes(rows, V)
    ws = _workspace(dev, ws_bytes)
    ws_bytes = ws.numel()
    with torch.cuda.device(dev):
        rc = lib.specdec_sample_rows(_ptr(x), _dtype_code(x), rows, V, stride, float(temperature), int(top_k),
                                     float(top_p), int(sample_mode), _ptr(u), seed & 0xFFFFFFFFFFFFFFFF,
                                     offset & 0xFFFFFFFFFFFFFFFF, seq_id0, lane_id, _ptr(tok), _ptr(ptok), _ptr(ws),
                                     ws_bytes, _stream())
    L.check(rc, "specdec_sample_rows")
    return [tok, ptok]


@sample_rows_op.register_fake
def _(logits, u, seed, offset, seq_id0, lane_id, temperature, top_k, top_p, sample_mode):
    rows = logits.numel() // logits.shape[-1]
    return [logits.new_empty(rows, dtype=torch.int64), logits.new_empty(rows, dtype=torch.float32)]


@torch.library.custom_op("specdec::sample_probs", mutates_args=())
def sample_probs_op(probs: Tensor, u: Optional[Tensor], sample_mode: int) -> Tensor:
    _need_cuda(probs, u)
    V = probs.shape[-1]
    p = probs.to(torch.float32).reshape(-1, V).contiguous()
    rows = p.shape[0]
    if u is not None:
        u = u.to(device=p.device, dtype=torch.float32).reshape(rows).contiguous()
    tok = torch.empty(rows, dtype=torch.int64, device=p.device)
    with torch.cuda.device(p.device):
        rc = L.lib().specdec_sample_probs(_ptr(p), rows, V, int(sample_mode), _ptr(u), _ptr(tok), _stream())
    L.check(rc, "specdec_sample_probs")
    return tok


@sample_probs_op.register_fake
def _(probs, u, sample_mode):
    return probs.new_empty(probs.numel() // probs.shape[-1], dtype=torch.int64)


def philox_uniform_op(seed: int, offset: int, seq_id0: int, B: int, gamma: int, device: torch.device) -> List[Tensor]:
    """The exact uniforms specdec::verify draws when u_accept / u_sample are None."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("specdec ops run on CUDA tensors only (no CPU fallback)")
    ua = torch.empty((B, gamma), dtype=torch.float32, device=device)
    us = torch.empty(B, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = L.lib().specdec_philox_uniform(seed & 0xFFFFFFFFFFFFFFFF, offset & 0xFFFFFFFFFFFFFFFF, seq_id0, B, gamma,
                                            _ptr(ua), _ptr(us), _stream())
    L.check(rc, "specdec_philox_uniform")
    return [ua, us]


@torch.library.custom_op("specdec::prune_kv", mutates_args={"tensors", "seq_lens"})
def prune_kv_op(tensors: List[Tensor], seq_lens: Tensor, discard: Tensor, zero_fill: bool, ptrs: Optional[Tensor]) -> None:
    """Per-sequence KV rollback on static [B,H,S_max,D] cache tensors (utils/caching.py:27-55 generalised).
    ptrs: optional cached device table of the tensors' addresses (int64 [n]); without zero_fill only the
    length vector changes and the tensors are not touched at all."""
    _prune_kv_impl(tensors, seq_lens, discard, zero_fill, ptrs)


def _prune_kv_impl(tensors, seq_lens: Tensor, discard: Tensor, zero_fill: bool, ptrs: Optional[Tensor]) -> None:
    _need_cuda(seq_lens, discard, *(tensors if zero_fill else ()))
    if not tensors:
        return
    B, H, S, D = tensors[0].shape
    if seq_lens.dtype != torch.int32 or not seq_lens.is_contiguous():
        raise ValueError("prune_kv: seq_lens must be a contiguous int32 tensor (updated in place)")
    dev = seq_lens.device
    if discard.dtype != torch.int32 or discard.device != dev or not discard.is_contiguous():
        discard = discard.to(device=dev, dtype=torch.int32).contiguous()
    n = 0
    if zero_fill:
        for t in tensors:
            if t.shape != (B, H, S, D) or not t.is_contiguous() or t.dtype != tensors[0].dtype:
                raise ValueError("prune_kv: all cache tensors must be contiguous [B,H,S_max,D] of one dtype")
        if ptrs is None:
            ptrs = kv_pointer_table(tensors)
        n = len(tensors)
    with torch.cuda.device(dev):
        rc = L.lib().specdec_prune_kv(_ptr(ptrs) if n else None, n, B, H, S, D, tensors[0].element_size(),
                                      _ptr(seq_lens), _ptr(discard), 1 if zero_fill else 0, _stream())
    L.check(rc, "specdec_prune_kv")


def kv_pointer_table(tensors) -> Tensor:
    """Device array of the cache tensors' addresses (build once per cache: StaticKVCache caches it)."""
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=tensors[0].device)


def topk_ids(logits: Tensor, k: int) -> Tensor:
    """[..., V] logits -> [..., k] int64 ids of the k largest (value desc, index asc): the n-gram filler tokens
    (ngram_assisted/ngram_assisted.py:149-155) without a V-wide probability row."""
    _need_cuda(logits)
    x, rows, V, stride = _rows_view(logits)
    out = torch.empty((rows, int(k)), dtype=torch.int64, device=logits.device)
    with torch.cuda.device(logits.device):
        rc = L.lib().specdec_topk_ids(_ptr(x), _dtype_code(x), rows, V, stride, int(k), _ptr(out), _stream())
    L.check(rc, "specdec_topk_ids")
    return out.reshape(*logits.shape[:-1], int(k))


def batch_writeback(res, generated: Tensor, step, g: int, finished: Tensor, n_acc: Tensor, end_tokens: Optional[Tensor],
                    n_active: Optional[Tensor] = None) -> None:
    """engine/infer_engine.py:300-336 on device-resident state (see include/specdec_b200.h).  generated [B, G] int64
    (row stride arbitrary), finished [B] uint8/bool, n_acc [B] int64, step: int or 1-element int64 CUDA tensor."""
    _need_cuda(generated, finished, n_acc, end_tokens, n_active)
    B = generated.shape[0]
    if generated.dtype != torch.int64 or generated.stride(1) != 1 or n_acc.dtype != torch.int64:
        raise ValueError("batch_writeback: generated / n_acc must be int64, generated contiguous along the step axis")
    fin = finished.view(torch.uint8) if finished.dtype == torch.bool else finished
    n_end = 0 if end_tokens is None else end_tokens.numel()
    step_dev, step_host = (step.data_ptr(), 0) if isinstance(step, Tensor) else (None, int(step))
    with torch.cuda.device(generated.device):
        rc = L.lib().specdec_batch_writeback(B, int(g), _ptr(res.n_accepted), _ptr(res.first_stop), _ptr(res.next_token),
                                             _ptr(generated), generated.stride(0), step_dev, step_host, _ptr(fin),
                                             _ptr(n_acc), _ptr(end_tokens) if n_end else None, n_end, _ptr(n_active),
                                             _stream())
    L.check(rc, "specdec_batch_writeback")


# ---- convenient python wrappers -------------------------------------------------------------
class VerifyResult:
    """Outputs of one verify step.  The eight tensors are views of ONE device buffer, created on first access."""
    __slots__ = ("_buf", "_B", "_g", "_v")
    _NAMES = ("n_accepted", "next_token", "accept_mask", "p_tok", "q_tok", "first_stop", "next_prob", "packed")

    def __init__(self, buf, B, g):
        self._buf, self._B, self._g, self._v = buf, B, g, None

    def _views(self):
        if self._v is None:
            self._v = _verify_views(self._buf, self._B, self._g)
        return self._v

    n_accepted = property(lambda self: self._views()[0])
    next_token = property(lambda self: self._views()[1])
    accept_mask = property(lambda self: self._views()[2])
    p_tok = property(lambda self: self._views()[3])
    q_tok = property(lambda self: self._views()[4])
    first_stop = property(lambda self: self._views()[5])
    next_prob = property(lambda self: self._views()[6])
    packed = property(lambda self: self._views()[7])

    def host(self):
        """ONE device->host copy (one sync) of everything a decode loop reads per step:
        -> (n_accepted [B], next_token [B], first_stop [B]) as python lists.  The reference pays three `.item()` syncs
        here (sampling/speculative_decoding.py:141-152, 172-187)."""
        B, g = self._B, self._g
        h = self._buf.cpu()
        n32 = _n32(B, g)
        i32 = h[:n32 * 4].view(torch.int32)
        nxt = h[n32 * 4:n32 * 4 + B * 8].view(torch.int64)
        return i32[0:B].tolist(), nxt.tolist(), i32[B:2 * B].tolist()


def _device_offset(offset, dev):
    """offset given as a 1-element int64 CUDA tensor -> its address (the kernels read the word when they run)."""
    if offset.device != dev or offset.dtype != torch.int64 or offset.numel() != 1:
        raise ValueError("a device-resident Philox offset must be a 1-element int64 tensor on the logits' device")
    return offset.data_ptr()


def fused_verify(target_logits, draft_logits, draft_tokens, u_accept=None, u_sample=None, *, seed=0, offset=0,
                 seq_id0=0, temperature=1.0, top_k=0, top_p=1.0, greedy=False, flags=0, stop_tokens=None) -> VerifyResult:
    """offset: int, or a 1-element int64 CUDA tensor (device-resident step counter: a captured CUDA graph of the step
    draws fresh uniforms on every replay once the caller bumps the tensor, e.g. `offset.add_(1)` inside the graph)."""
    if isinstance(offset, Tensor):
        offset, flags = _device_offset(offset, target_logits.device), int(flags) | L.OFFSET_DEVICE
    if stop_tokens is not None and not isinstance(stop_tokens, Tensor):
        stop_tokens = torch.as_tensor(list(stop_tokens), dtype=torch.int64, device=target_logits.device)
    if stop_tokens is not None and stop_tokens.numel() == 0:
        stop_tokens = None
    # direct call of the implementation (same code the registered torch op `specdec::verify` runs) -- skips
    # the dispatcher's per-call overhead, which matters at small batch
    return VerifyResult(*_verify_impl(target_logits, draft_logits, draft_tokens, u_accept, u_sample, int(seed), int(offset),
                                      int(seq_id0), float(temperature), int(top_k), float(top_p),
                                      L.SAMPLE_GREEDY if greedy else L.SAMPLE_INVCDF, int(flags), stop_tokens))


class GraphedVerify:
    """One verify step captured ONCE into a CUDA graph and replayed: the launch cost of a step drops from four
    enqueues plus Python argument handling to a single cudaGraphLaunch, which is what bounds the step at small batch
    (sampling/speculative_decoding.py is batch 1).  The inputs are static buffers (`target`, `draft`, `tokens`: pass
    your model's output buffers to alias them, or copy into the ones allocated here); the Philox offset is a device
    word the graph itself bumps, so every replay draws fresh uniforms (SPECDEC_OFFSET_DEVICE)."""

    def __init__(self, target: Tensor, draft: Optional[Tensor], tokens: Tensor, *, seed=0, offset0=0, seq_id0=0,
                 temperature=1.0, top_k=0, top_p=1.0, greedy=False, flags=0, stop_tokens=None):
        _need_cuda(target, draft, tokens)
        self.target, self.draft, self.tokens = target, draft, tokens
        dev = target.device
        self.offset = torch.full((1,), int(offset0), dtype=torch.int64, device=dev)
        kw = dict(seed=seed, offset=self.offset, seq_id0=seq_id0, temperature=temperature, top_k=top_k, top_p=top_p,
                  greedy=greedy, flags=flags, stop_tokens=stop_tokens)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture: lazy per-device state, workspace growth
            fused_verify(target, draft, tokens, None, None, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fused_verify(target, draft, tokens, None, None, **kw)
            self.offset.add_(1)

    def __call__(self) -> VerifyResult:
        """Replays the step on the buffers' CURRENT contents; the returned VerifyResult is overwritten by the next call."""
        self.graph.replay()
        return self.result


def process_probs(logits, temperature=1.0, top_k=0, top_p=1.0):
    return process_probs_op(logits, float(temperature), int(top_k), float(top_p))


def sample_rows(logits, u=None, *, seed=0, offset=0, seq_id0=0, lane_id=0, temperature=1.0, top_k=0, top_p=1.0,
                greedy=False):
    if isinstance(offset, Tensor):
        offset, lane_id = _device_offset(offset, logits.device), int(lane_id) | L.LANE_OFFSET_DEVICE
    return sample_rows_op(logits, u, int(seed), int(offset), int(seq_id0), int(lane_id), float(temperature),
                          int(top_k), float(top_p), L.SAMPLE_GREEDY if greedy else L.SAMPLE_INVCDF)


def sample_probs(probs, u=None, greedy=False):
    return sample_probs_op(probs, u, L.SAMPLE_GREEDY if greedy else L.SAMPLE_INVCDF)


def philox_uniform(seed, offset, seq_id0, B, gamma, device="cuda"):
    return philox_uniform_op(int(seed), int(offset), int(seq_id0), int(B), int(gamma), torch.device(device))


def prune_kv(tensors, seq_lens, discard, zero_fill=False, ptrs=None):
    """zero_fill=False (default): only the length vector is updated -- the valid prefix [0, len_b) is the
    reference's pruned view; zero_fill=True also clears the discarded positions."""
    # direct call of the implementation the registered op `specdec::prune_kv` runs (skips ~20 us of dispatcher
    # overhead per call for a ~3 us kernel)
    _prune_kv_impl(tensors, seq_lens, discard, bool(zero_fill), ptrs)
    return seq_lens
