"""Uniform sources for the verify path.

Production: `PhiloxUniforms` -- nothing is drawn on the host; the kernels generate
Philox4x32-10 uniforms keyed by (seed, step offset, GLOBAL sequence id, position, stream), so the
results do not depend on batch sharding or launch geometry (RNG contract).
Tests: `InjectedUniforms` replays given streams in the order the reference consumes
torch.rand / sample() (sampling/speculative_decoding.py:93,123,139,171).
"""
from __future__ import annotations

import numpy as np
import torch


class PhiloxUniforms:
    injected = False

    def __init__(self, seed: int | None = None):
        self.seed = int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self.offset = 0

    def next_offset(self) -> int:
        o = self.offset
        self.offset += 1
        return o


_DEFAULT = None


def default_uniforms() -> PhiloxUniforms:
    """The process-wide generator every API call without an explicit `uniforms` draws from: keyed by
    torch.initial_seed() (re-keyed, offset 0, when the user reseeds with torch.manual_seed) and with an offset that
    PERSISTS across calls -- like the reference's global torch generator, successive speculative_generate /
    LogitsProcessor.sample / batch calls never replay one another's uniforms.  Every consumer takes a fresh offset per
    kernel call, so two calls never share a (seed, offset) pair whatever Philox lane they read."""
    global _DEFAULT
    seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
    if _DEFAULT is None or _DEFAULT.seed != seed:
        _DEFAULT = PhiloxUniforms(seed)
    return _DEFAULT


class InjectedUniforms:
    """sample_u feeds every sample() call, accept_u feeds the rand(gamma) of the accept test."""
    injected = True

    def __init__(self, sample_u, accept_u, device="cuda"):
        self.s = np.asarray(sample_u, dtype=np.float32).reshape(-1)
        self.a = np.asarray(accept_u, dtype=np.float32).reshape(-1)
        self.si = 0
        self.ai = 0
        self.device = device
        self.seed = 0
        self.offset = 0

    def next_offset(self) -> int:
        o = self.offset
        self.offset += 1
        return o

    def _take(self, arr, i, n):
        if i + n > arr.size:
            raise RuntimeError("injected uniform stream exhausted")
        return torch.from_numpy(arr[i:i + n].copy()).to(self.device)

    def sample(self, n=1):
        t = self._take(self.s, self.si, n)
        self.si += n
        return t

    def accept(self, n):
        t = self._take(self.a, self.ai, n)
        self.ai += n
        return t

    def peek_accept(self, n):
        return self._take(self.a, self.ai, n)

    def peek_sample(self, n):
        return self._take(self.s, self.si, n)

    def skip_sample(self, n):
        self.si += n

    def skip_accept(self, n):
        self.ai += n
