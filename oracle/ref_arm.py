"""The UNMODIFIED reference verify step on CPU, driven through its own public API (TEST / BASELINE INFRASTRUCTURE).

`speculative_generate` (oracle/_ref/sampling/speculative_decoding.py, copied verbatim from the reference by
oracle/make_ref.py) is called once per sequence with two replay models that hand it the synthetic logit rows, so that
exactly ONE speculative step runs through the reference's stock code path: gamma x (LogitsProcessor.__call__ + sample)
on the drafter rows (:110-124), the processor on the gamma target rows (:135-136), torch.rand + p/q + the
first-rejection loop (:139-146), the stop scan (:151), max_fn(p - q) or the bonus row (:160-171) and the final sample
(:172).  The step ends when the drafter is asked for a second round (a private exception, caught here).

Only bench.py's `--impl reference` / `cpu_baseline` legs and tests may import this module.
"""
import hashlib
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    mf = os.path.join(REF, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    try:
        for rel, h in json.load(open(mf))["files"].items():
            if hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest() != h:
                return False
    except Exception:
        return False
    return True


def modules():
    if "termcolor" not in sys.modules:  # utils/printing.py imports it for debug prints only
        stub = types.ModuleType("termcolor")
        stub.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = stub
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import utils.logits_processor as lp
    import sampling.speculative_decoding as sd
    assert os.path.abspath(sd.__file__).startswith(REF), sd.__file__
    return lp, sd


class _StepDone(Exception):
    pass


class _Out:
    def __init__(self, logits):
        self.logits, self.past_key_values = logits, None


class _Cfg:
    def __init__(self, V, L):
        self.vocab_size, self.max_position_embeddings = V, L


class _Drafter:
    """call k of the step returns drafter row k as the last position's logits; a (gamma+1)-th call ends the step"""

    def __init__(self, rows, V):
        self.rows, self.k, self.device, self.config = rows, 0, torch.device("cpu"), _Cfg(V, 1 << 20)

    def __call__(self, input_ids=None, past_key_values=None, use_cache=False):
        if self.k >= self.rows.shape[0]:
            raise _StepDone()
        r = self.rows[self.k].reshape(1, 1, -1)
        self.k += 1
        return _Out(r)


class _Target:
    def __init__(self, rows, V):
        self.rows, self.calls, self.device, self.config = rows, 0, torch.device("cpu"), _Cfg(V, 1 << 20)

    def __call__(self, input_ids=None, past_key_values=None, use_cache=False):
        if self.calls:  # (a rejection at the last position leaves a zero-draft second round: not part of this step)
            raise _StepDone()
        self.calls += 1
        return _Out(self.rows.unsqueeze(0))  # [1, gamma+1, V]: positions current-1 .. current+gamma-1 with prompt_len 1


def make_processor(mode):
    lp, _ = modules()
    if mode["greedy"]:
        return lp.GreedyProcessor(mode["temperature"])
    k, p = mode["top_k"], mode["top_p"]
    if k and p < 1.0:
        return lp.TopKNucleusProcessor(mode["temperature"], k, p)
    if k:
        return lp.TopKProcessor(mode["temperature"], k)
    if p < 1.0:
        return lp.NucleusProcessor(mode["temperature"], p)
    return lp.MultinomialProcessor(mode["temperature"])


def verify_step(target_rows, draft_rows, mode, processor=None):
    """One reference speculative step for one sequence: target_rows [gamma+1, V], draft_rows [gamma, V] (fp32 CPU)."""
    _, sd = modules()
    g, V = draft_rows.shape
    proc = processor or make_processor(mode)
    try:
        sd.speculative_generate([0], _Drafter(draft_rows, V), _Target(target_rows, V), gamma=g, logits_processor=proc,
                                max_gen_len=g + 1, eos_tokens_id=-1, pad_token_id=0, use_cache=False,
                                first_target=False)
    except _StepDone:
        pass
