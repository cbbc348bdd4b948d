"""Drop-in n-gram storages (ngram_assisted/ngram_storage.py:5-249) backed by device hash tables
(csrc/ngram.cu).  Same interface: next_token / has_gram / update / initialize / reset.

Differences by design:
  * tables live in HBM; `next_token` / `update` / `initialize` are one kernel launch for the batch;
  * `table_ids` (optional) gives every sequence its own logical table -- the layout batched
    n-gram-assisted decoding needs; default = one shared table, as in the reference;
  * `lookup_chain` runs the gamma chained next_token() calls of ngram_assisted.py:95-99 in one launch;
  * unknown contexts take a caller-supplied fallback token or one drawn ON THE DEVICE (Philox keyed by the storage's
    seed, the call number, the sequence and the position) instead of a host torch.randint
    (ngram_storage.py:84,165); an empty table is a miss, not a KeyError (:174).
"""
from __future__ import annotations

import abc
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L


class INgramStorage(abc.ABC):
    """Interface of Ngram-Storage (ngram_assisted/ngram_storage.py:5-69)."""

    def __init__(self, n: int, vocab_size: int):
        assert n > 1, "n should be greater than 1"
        self.n = n
        self.vocab_size = vocab_size

    @abc.abstractmethod
    def next_token(self, input_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @abc.abstractmethod
    def has_gram(self, ngram: torch.Tensor) -> bool: ...

    @abc.abstractmethod
    def update(self, input_ids: torch.Tensor, next_tokens: torch.Tensor): ...

    @abc.abstractmethod
    def initialize(self, input_ids: torch.Tensor): ...

    @abc.abstractmethod
    def reset(self): ...


class _DeviceNGram(INgramStorage):
    _one_level = 0

    def __init__(self, n: int, vocab_size: int, n_tables: int = 1, grams_per_table: int = 1 << 16,
                 counts_per_table: int = 1 << 18, device="cuda", seed: int = 0):
        super().__init__(n, vocab_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("device n-gram tables need a CUDA device (no CPU fallback)")
        if n - 1 > 8:
            raise ValueError("device n-gram tables hold contexts of at most 8 tokens (n <= 9)")
        self.n_tables = n_tables
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().specdec_ngram_create(C.byref(self._h), n, vocab_size, n_tables, grams_per_table,
                                                 counts_per_table, self._one_level), "specdec_ngram_create")
            L.check(L.lib().specdec_ngram_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF), "specdec_ngram_seed")

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                L.lib().specdec_ngram_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # -- helpers
    def _prep(self, input_ids, lens, table_ids):
        ids = input_ids if input_ids.dim() == 2 else input_ids.reshape(1, -1)
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous()
        B, ml = ids.shape
        if lens is None:
            key = (B, ml)  # (full rows: the length vector is cached per shape)
            cl = getattr(self, "_full_lens", None)
            if cl is None or cl[0] != key:
                self._full_lens = cl = (key, torch.full((B,), ml, dtype=torch.int32, device=self.device))
            lens = cl[1]
        else:
            lens = lens.to(device=self.device, dtype=torch.int32).contiguous()
        if table_ids is not None:
            # the kernels index the tables with these ids unchecked: validate once per distinct tensor (one sync),
            # not per call
            key = (table_ids.data_ptr(), table_ids.numel(), table_ids._version)
            if key != getattr(self, "_tab_ok", None):
                if table_ids.numel() != B or (table_ids.numel() and (int(table_ids.min()) < 0 or int(table_ids.max()) >= self.n_tables)):
                    raise ValueError(f"table_ids must hold one id in [0, {self.n_tables}) per sequence")
                self._tab_ok = key
            table_ids = table_ids.to(device=self.device, dtype=torch.int32).contiguous()
        return ids, lens, table_ids, B, ml

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- INgramStorage
    def lookup_chain(self, input_ids, gamma: int, lens=None, table_ids=None, fallback=None):
        ids, lens, table_ids, B, ml = self._prep(input_ids, lens, table_ids)
        if fallback is not None:  # None: drawn on the device (no host RNG, no copy)
            fallback = fallback.to(device=self.device, dtype=torch.int64).reshape(B, gamma).contiguous()
        drafts = torch.empty((B, gamma), dtype=torch.int64, device=self.device)
        known = torch.empty((B, gamma), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().specdec_ngram_lookup_chain(self._h, self._p(ids), self._p(lens), self._p(table_ids), B, ml,
                                                       gamma, self._p(fallback), self._p(drafts), self._p(known),
                                                       self._stream()), "specdec_ngram_lookup_chain")
        return drafts, known.view(torch.bool)  # (0/1 bytes written by the kernel: a view, not a conversion launch)

    def next_token(self, input_ids, lens=None, table_ids=None, fallback=None):
        d, k = self.lookup_chain(input_ids, 1, lens, table_ids, fallback)
        return d[:, 0], k[:, 0]

    def update(self, input_ids, next_tokens, lens=None, table_ids=None):
        ids, lens, table_ids, B, ml = self._prep(input_ids, lens, table_ids)
        nt = next_tokens.to(device=self.device, dtype=torch.int64).reshape(B, -1).contiguous()
        with torch.cuda.device(self.device):
            L.check(L.lib().specdec_ngram_update(self._h, self._p(ids), self._p(lens), self._p(table_ids), B, ml,
                                                 self._p(nt), nt.shape[1], self._stream()), "specdec_ngram_update")

    def update_chain(self, ids_row, start: int, accepted, fillers=None, table_id: int = 0):
        """The per-position updates of ONE n-gram-assisted step (ngram_assisted/ngram_assisted.py:149-155) as one launch:
            for i in range(len(accepted)):
                update(ids_row[:start + i], [accepted[i]]);  update(ids_row[:start + i], fillers[i])   # if fillers
        The rows are applied in exactly this order by the table's thread (the arg-max rule of the reference depends on
        the order of the updates); rows with fewer tokens are padded with -1, which the kernel skips.
        accepted: int64 [n1] (device), fillers: int64 [>= n1, k] (device) or None."""
        n1 = int(accepted.numel())
        if n1 == 0:
            return
        dev = self.device
        k = int(fillers.shape[1]) if fillers is not None else 1
        rows = n1 * (2 if fillers is not None else 1)
        nt = torch.full((rows, k), -1, dtype=torch.int64, device=dev)
        pos = torch.arange(start, start + n1, dtype=torch.int32, device=dev)
        if fillers is not None:
            nt[0::2, 0] = accepted.to(dev)
            nt[1::2] = fillers[:n1].to(dev)
            lens = pos.repeat_interleave(2)
        else:
            nt[:, 0] = accepted.to(dev)
            lens = pos
        ids = ids_row.reshape(1, -1)[:, :start + n1].expand(rows, -1)
        tabs = None if self.n_tables == 1 else torch.full((rows,), table_id, dtype=torch.int32, device=dev)
        self.update(ids, nt, lens=lens, table_ids=tabs)

    def initialize(self, input_ids, lens=None, table_ids=None):
        ids, lens, table_ids, B, ml = self._prep(input_ids, lens, table_ids)
        with torch.cuda.device(self.device):
            L.check(L.lib().specdec_ngram_initialize(self._h, self._p(ids), self._p(lens), self._p(table_ids), B, ml,
                                                     self._stream()), "specdec_ngram_initialize")

    def reset(self):
        with torch.cuda.device(self.device):
            L.check(L.lib().specdec_ngram_reset(self._h, self._stream()), "specdec_ngram_reset")

    def has_gram(self, ngram: torch.Tensor, table_id: int = 0) -> bool:
        """ngram_storage.py:98-106 / :181-193, exact (a probe of the gram and count tables, no insert): True iff the
        final token of `ngram` was ever counted after the context made of its last j tokens."""
        if ngram.numel() < 1:
            return False
        tabs = None if self.n_tables == 1 else torch.tensor([table_id], dtype=torch.int32, device=self.device)
        ids, lens, tabs, B, ml = self._prep(ngram.reshape(1, -1), None, tabs)
        out = torch.empty(1, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().specdec_ngram_has_gram(self._h, self._p(ids), self._p(lens), self._p(tabs), 1, ml, self._p(out),
                                                   self._stream()), "specdec_ngram_has_gram")
        return bool(out[0])

    def status(self):
        out = (C.c_int32 * 2)()
        L.check(L.lib().specdec_ngram_status(self._h, out), "specdec_ngram_status")
        return {"overflow": bool(out[0]), "grams_used_max": int(out[1])}


class NGramStorage(_DeviceNGram):
    """Multi-level storage: context lengths j in [2, n-1], longest first (ngram_storage.py:154-249)."""
    _one_level = 0


class OneLevelNGramStorage(_DeviceNGram):
    """Single context length n-1 (ngram_storage.py:73-150)."""
    _one_level = 1
