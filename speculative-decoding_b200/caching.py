"""KV-cache rollback: drop-in prune_cache (utils/caching.py:6-77) + the per-sequence static cache
the batched path needs.

* tuple caches: zero-copy views exactly like the reference (utils/caching.py:27-55) -- nothing to
  accelerate, a view is already free.
* DynamicCache: the reference touches cache.key_cache / value_cache / _seen_tokens (:72-75), which
  current transformers no longer has; here the crop() API is used when present, else those fields.
* StaticKVCache: [B,H,S_max,D] tensors + int32 lengths; rollback with a DIFFERENT discard count per
  sequence runs the specdec::prune_kv CUDA kernel (no reference implementation exists: the
  reference's batched loop never prunes, engine/infer_engine.py:239-243).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple, Union

import torch
from torch import Tensor

from . import ops

try:  # transformers is only needed for the DynamicCache branch
    from transformers.cache_utils import DynamicCache
except Exception:  # pragma: no cover
    DynamicCache = None


def prune_cache(cache, num_tokens_to_discard: int):
    """Prune the cache by removing the specified number of tokens from the end."""
    if cache is None:
        return None
    if isinstance(cache, tuple):
        return prune_tuple_cache(cache, num_tokens_to_discard)
    if isinstance(cache, StaticKVCache):
        return cache.rollback(num_tokens_to_discard)
    if DynamicCache is not None and isinstance(cache, DynamicCache):
        return prune_dynamic_cache(cache, num_tokens_to_discard)
    raise ValueError("Unsupported cache type.")


def prune_tuple_cache(cache: Tuple[Tuple[Tensor, Tensor]], num_tokens_to_discard: int):
    if cache is None:
        return None
    new_cache = []
    for layer_cache in cache:
        if layer_cache is None:
            new_cache.append(None)
            continue
        new_cache.append(tuple(t[:, :, :-num_tokens_to_discard, :] for t in layer_cache))
    return tuple(new_cache)


def prune_dynamic_cache(cache, num_tokens_to_discard: int):
    if cache is None:
        return None
    if hasattr(cache, "key_cache"):  # transformers < 4.54 layout used by the reference
        for layer in range(len(cache)):
            cache.key_cache[layer] = cache.key_cache[layer][:, :, :-num_tokens_to_discard, :]
            cache.value_cache[layer] = cache.value_cache[layer][:, :, :-num_tokens_to_discard, :]
        cache._seen_tokens -= num_tokens_to_discard
        return cache
    cache.crop(cache.get_seq_length() - num_tokens_to_discard)
    return cache


class StaticKVCache:
    """Per-sequence-length KV cache: `tensors` = flat list of [B,H,S_max,D] tensors (K and V of every
    layer), `seq_lens` int32 [B]."""

    def __init__(self, tensors: Sequence[Tensor], seq_lens: Tensor):
        self.tensors: List[Tensor] = list(tensors)
        self.seq_lens = seq_lens.to(torch.int32).contiguous()
        self._ptrs = None  # device table of the tensors' addresses, built on the first zero-filling rollback

    def rollback(self, discard: Union[int, Tensor], zero_fill: bool = False) -> "StaticKVCache":
        """Drops the last discard[b] positions of sequence b.  The length vector IS the rollback (one ~3 us launch);
        zero_fill=True additionally clears the discarded positions."""
        B = self.seq_lens.shape[0]
        if not isinstance(discard, Tensor):
            discard = torch.full((B,), int(discard), dtype=torch.int32, device=self.seq_lens.device)
        if zero_fill and self._ptrs is None:
            self._ptrs = ops.kv_pointer_table(self.tensors)
        ops.prune_kv(self.tensors, self.seq_lens, discard, zero_fill, self._ptrs if zero_fill else None)
        return self

    def as_tuple_views(self, b: int):
        """The reference-shaped tuple cache of sequence b: ((K,V),...) views [1,H,len_b,D]."""
        n = int(self.seq_lens[b])
        ts = [t[b:b + 1, :, :n, :] for t in self.tensors]
        return tuple((ts[i], ts[i + 1]) for i in range(0, len(ts), 2))
