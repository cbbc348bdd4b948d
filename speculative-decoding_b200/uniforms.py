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
