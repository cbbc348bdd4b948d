"""Drives the UNMODIFIED reference (/root/reference) on CPU to pin the oracle.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does
not exist on the GPU box, so nothing run there (pytest -m gpu, smoke(), bench.py) may
import this module.  tests/golden/make_golden.py uses it to generate the committed
fixtures; tests/test_oracle_vs_reference.py re-runs it live when the reference is present.

Harness pieces:
  * termcolor stub so `sampling`, `ngram_assisted`, `engine` import.
  * `stable_sort()`: the two torch.sort calls of utils/logits_processor.py:74,96 get
    stable=True (tie order of equal logits is otherwise arbitrary).
  * `InjectedSampleMixin`: sample() = inverse CDF on an injected uniform instead of
    torch.multinomial (utils/logits_processor.py:48-49), same shapes/dtypes.
  * `patched_rand()`: torch.rand inside the reference modules pops from a given stream.
  * Fake models replaying position-indexed synthetic logits.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("SPECDEC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "utils"))


def _ensure_importable():
    if "termcolor" not in sys.modules:
        stub = types.ModuleType("termcolor")
        stub.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = stub
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def ref_modules():
    """-> (utils.logits_processor, sampling.speculative_decoding, ngram_assisted pkg, engine.infer_engine)"""
    _ensure_importable()
    import utils.logits_processor as lp
    import sampling.speculative_decoding as sd
    import ngram_assisted as ng
    import ngram_assisted.ngram_assisted as nga
    import engine.infer_engine as ie
    import utils.caching as caching
    return types.SimpleNamespace(lp=lp, sd=sd, ng=ng, nga=nga, ie=ie, caching=caching)


class _TorchProxy:
    """Forwards to torch, overriding selected attributes (used to patch one module's `torch` global)."""

    def __init__(self, **over):
        self.__dict__["_over"] = over

    def __getattr__(self, k):
        o = self.__dict__["_over"]
        return o[k] if k in o else getattr(torch, k)


@contextlib.contextmanager
def patched(module, **over):
    old = module.torch
    module.torch = _TorchProxy(**over)
    try:
        yield
    finally:
        module.torch = old


def _stable_sort(x, *a, **k):
    k.setdefault("stable", True)
    return torch.sort(x, *a, **k)


def stable_sort(lp_module):
    return patched(lp_module, sort=_stable_sort)


class UniformStream:
    """FIFO of pre-drawn uniforms; records how many were consumed."""

    def __init__(self, values):
        self.v = np.asarray(values, dtype=np.float32).reshape(-1)
        self.i = 0

    def pop(self, n=1):
        out = self.v[self.i:self.i + n]
        if out.size < n:
            raise RuntimeError("uniform stream exhausted")
        self.i += n
        return out


def inv_cdf_reference(probs: torch.Tensor, u: float) -> int:
    """Inverse CDF exactly as the oracle defines it, but on the REFERENCE's fp32 probs:
    float64 cumulative sums of the probabilities, first index with cum > u*total."""
    p = probs.detach().double().reshape(-1).clamp_min(0)
    cum = torch.cumsum(p, 0)
    tgt = float(u) * float(cum[-1])
    j = int(torch.searchsorted(cum, torch.tensor(tgt, dtype=torch.float64), right=True))
    nz = torch.nonzero(p > 0).reshape(-1)
    return min(j, int(nz[-1])) if nz.numel() else 0


def make_processor(kind: str, temperature=1.0, top_k=0, top_p=1.0, sample_stream: UniformStream | None = None):
    """Reference LogitsProcessor of the given kind whose sample() is inverse-CDF on `sample_stream`
    (greedy keeps the reference argmax)."""
    lp = ref_modules().lp
    base = dict(greedy=lp.GreedyProcessor, multinomial=lp.MultinomialProcessor, topk=lp.TopKProcessor,
                nucleus=lp.NucleusProcessor, topk_nucleus=lp.TopKNucleusProcessor)[kind]

    class Injected(base):
        def sample(self, probs):
            if kind == "greedy":
                return super().sample(probs)
            shp = probs.shape[:-1]
            flat = probs.reshape(-1, probs.shape[-1])
            out = torch.tensor([inv_cdf_reference(flat[r], sample_stream.pop(1)[0]) for r in range(flat.shape[0])],
                               dtype=torch.long)
            return out.reshape(*shp, 1) if len(shp) else out.reshape(1)

    if kind == "greedy":
        return Injected(temperature)
    if kind == "multinomial":
        return Injected(temperature)
    if kind == "topk":
        return Injected(temperature, top_k)
    if kind == "nucleus":
        return Injected(temperature, top_p)
    return Injected(temperature, top_k, top_p)


class _Cfg:
    def __init__(self, V, max_pos):
        self.vocab_size = V
        self.max_position_embeddings = max_pos


class _Out:
    def __init__(self, logits, pkv=None):
        self.logits = logits
        self.past_key_values = pkv


class PositionTableModel:
    """Decoder-only fake: logits at position t are table[t] regardless of the tokens
    (fields used by the reference: sampling/speculative_decoding.py:69,73,77,86-92,113-120,129-135)."""

    def __init__(self, table: torch.Tensor, max_pos=None):
        self.table = table  # [L, V]
        self.device = torch.device("cpu")
        self.config = _Cfg(table.shape[-1], max_pos or table.shape[0])

    def __call__(self, input_ids=None, past_key_values=None, use_cache=False, **kw):
        L = input_ids.shape[-1]
        # clone: TopKProcessor mutates its argument in place (utils/logits_processor.py:62)
        return _Out(self.table[:L].clone().unsqueeze(0), None)


def run_speculative_generate(prompt, q_table, p_table, kind, *, gamma, max_gen_len, temperature=1.0, top_k=0,
                             top_p=1.0, eos=-1, pad=0, skip_sample_adjustment=False, first_target=True,
                             sample_u=None, accept_u=None):
    """Runs the reference's speculative_generate (sampling/speculative_decoding.py:23) unmodified.
    Returns (tokens, accept_rate, n_sample_u_used, n_accept_u_used)."""
    mods = ref_modules()
    s_stream = UniformStream(sample_u)
    a_stream = UniformStream(accept_u)
    proc = make_processor(kind, temperature, top_k, top_p, s_stream)

    def fake_rand(n, device=None, **k):
        return torch.from_numpy(a_stream.pop(int(n)).copy())

    with stable_sort(mods.lp), patched(mods.sd, rand=fake_rand):
        toks, rate = mods.sd.speculative_generate(
            list(prompt), PositionTableModel(q_table), PositionTableModel(p_table), tokenizer=None, gamma=gamma,
            logits_processor=proc, max_gen_len=max_gen_len, eos_tokens_id=eos, pad_token_id=pad, use_cache=False,
            skip_sample_adjustment=skip_sample_adjustment, first_target=first_target, debug=False)
    return toks, rate, s_stream.i, a_stream.i


def run_ngram_generate(prompt, p_table, kind, *, ngram_n, gamma, max_gen_len, filler_top_k=3, temperature=1.0,
                       top_k=0, top_p=1.0, eos=-1, pad=0, stop_if_unknown=False, sample_u=None, fallback_tokens=None):
    """Runs ngram_assisted_speculative_generate (ngram_assisted/ngram_assisted.py:11) unmodified with the
    reference NGramStorage; the storage's random fallback (torch.randint, ngram_storage.py:165) is
    replaced by the given deterministic fallback token stream."""
    mods = ref_modules()
    s_stream = UniformStream(sample_u)
    proc = make_processor(kind, temperature, top_k, top_p, s_stream)
    V = p_table.shape[-1]
    storage = mods.ng.NGramStorage(ngram_n, V)
    fb = {"i": 0}

    def fake_randint(high, size=None, **k):
        n = int(size[0])
        out = torch.tensor([int(fallback_tokens[(fb["i"] + t) % len(fallback_tokens)]) for t in range(n)], dtype=torch.long)
        fb["i"] += n
        return out

    import ngram_assisted.ngram_storage as ngs
    with stable_sort(mods.lp), patched(ngs, randint=fake_randint):
        toks, rate = mods.nga.ngram_assisted_speculative_generate(
            list(prompt), storage, PositionTableModel(p_table), tokenizer=None, gamma=gamma,
            filler_top_k=filler_top_k, logits_processor=proc, max_gen_len=max_gen_len, eos_tokens_id=eos,
            pad_token_id=pad, use_cache=False, first_target=True, stop_if_unknown=stop_if_unknown, debug=False)
    return toks, rate, s_stream.i, fb["i"]


class BatchTableDrafter:
    """Batched fake drafter for engine/infer_engine.py:204-243: position tracked through past_key_values."""

    def __init__(self, table):  # [B, L, V]
        self.table = table
        self.device = torch.device("cpu")
        self.config = _Cfg(table.shape[-1], table.shape[1])

    def __call__(self, input_ids, attention_mask=None, past_key_values=None, use_cache=True, **kw):
        B, L = input_ids.shape
        start = 0 if past_key_values is None else past_key_values
        return _Out(self.table[:, start:start + L].clone(), start + L)


class BatchTableTarget:
    def __init__(self, table):
        self.table = table
        self.device = torch.device("cpu")
        self.config = _Cfg(table.shape[-1], table.shape[1])

    def __call__(self, input_ids, **kw):
        return _Out(self.table[:, :input_ids.shape[1]].clone(), None)


def run_batch_speculative_generate(input_ids, q_table, p_table, *, gamma, gen_len, end_tokens=(), sample_u=None,
                                   accept_u=None):
    """Runs engine/infer_engine.py:149 batch_speculative_generate unmodified; torch.multinomial (:246,321,325)
    -> inverse CDF on sample_u, torch.rand(1) (:305) -> accept_u."""
    mods = ref_modules()
    s_stream = UniformStream(sample_u)
    a_stream = UniformStream(accept_u)

    def fake_multinomial(probs, n, **k):
        flat = probs.reshape(-1, probs.shape[-1])
        out = torch.tensor([[inv_cdf_reference(flat[r], s_stream.pop(1)[0])] for r in range(flat.shape[0])],
                           dtype=torch.long)
        return out if probs.dim() > 1 else out.reshape(1)

    def fake_rand(n, device=None, **k):
        return torch.from_numpy(a_stream.pop(int(n)).copy())

    ctx = types.SimpleNamespace(drafter=BatchTableDrafter(q_table), target=BatchTableTarget(p_table), gamma=gamma,
                                gen_len=gen_len, end_tokens=list(end_tokens), target_device=torch.device("cpu"))
    B = input_ids.shape[0]
    with patched(mods.ie, multinomial=fake_multinomial, rand=fake_rand):
        outs, rates = mods.ie.batch_speculative_generate(ctx, input_ids, torch.ones_like(input_ids), B, None)
    return [o.tolist() for o in outs], rates, s_stream.i, a_stream.i
