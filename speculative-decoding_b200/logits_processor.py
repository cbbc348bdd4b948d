"""Drop-in LogitsProcessor family (same names / constructor arguments / call contract as
utils/logits_processor.py:7-103 of the reference) backed by the fused CUDA ops.

  __call__(logits[..., V]) -> probs[..., V]   (same dtype as logits; computed in fp32)
  _process(logits)         -> masked logits (-1e20 on removed tokens), NOT in place
                              (the reference's TopKProcessor mutates its argument, :62)
  sample(probs[..., V])    -> token ids [..., 1] int64
The fused decode loops never materialise probabilities: they read (temperature, top_k, top_p,
greedy) from the processor and call specdec::verify / specdec::sample_rows directly.
"""
from __future__ import annotations

import abc

import torch
from torch import Tensor

from . import ops
from .uniforms import PhiloxUniforms, default_uniforms  # noqa: F401


class LogitsProcessor(abc.ABC):
    """Logits processors for sampling."""

    top_k: int = 0
    top_p: float = 1.0
    greedy: bool = False

    def __init__(self, temperature: float):
        self.temperature = temperature
        self.uniforms = None  # None: the process-wide default generator (uniforms.default_uniforms)

    def __call__(self, logits: Tensor) -> Tensor:
        probs, _ = ops.process_probs(logits, self.temperature, self.top_k, self.top_p)
        return probs.to(logits.dtype)

    def _process(self, logits: Tensor) -> Tensor:
        if self.top_k <= 0 and not (0.0 < self.top_p < 1.0):
            return logits
        probs, _ = ops.process_probs(logits, self.temperature, self.top_k, self.top_p)
        return torch.where(probs > 0, logits, torch.full_like(logits, -1e20))

    def sample(self, probs: Tensor) -> Tensor:
        shp = probs.shape[:-1]
        if self.greedy:
            tok = ops.sample_probs(probs, None, greedy=True)
        else:
            rows = probs.numel() // probs.shape[-1]
            un = self.uniforms if self.uniforms is not None else default_uniforms()
            _, u = ops.philox_uniform(un.seed, un.next_offset(), 0, rows, 1, probs.device)  # the sample lane
            tok = ops.sample_probs(probs, u.reshape(-1), greedy=False)
        return tok.reshape(*shp, 1) if len(shp) else tok.reshape(1)

    # parameters consumed by the fused ops
    def fused_params(self):
        return dict(temperature=float(self.temperature), top_k=int(self.top_k), top_p=float(self.top_p),
                    greedy=bool(self.greedy))


class GreedyProcessor(LogitsProcessor):
    """Greedy: Most probable token."""
    greedy = True

    def __init__(self, temperature: float = 1):
        super().__init__(temperature)


class MultinomialProcessor(LogitsProcessor):
    """Multinomial: Random sampling."""

    def __init__(self, temperature: float):
        super().__init__(temperature)


class TopKProcessor(MultinomialProcessor):
    """Top-k: Top-k sampling."""

    def __init__(self, temperature: float, top_k: int):
        super().__init__(temperature)
        self.top_k = top_k


class NucleusProcessor(MultinomialProcessor):
    """Nucleus: Top-p sampling."""

    def __init__(self, temperature: float, top_p: float):
        super().__init__(temperature)
        self.top_p = top_p


class TopKNucleusProcessor(MultinomialProcessor):
    """Top-k and nucleus: Top-k sampling with top-p fallback."""

    def __init__(self, temperature: float, top_k: int, top_p: float):
        super().__init__(temperature)
        self.top_k = top_k
        self.top_p = top_p
