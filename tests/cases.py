"""Seeded synthetic inputs shared by the CPU and GPU tests (SURVEY.md 8d "concrete synthetic inputs")."""
import numpy as np
import torch

DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}

# (name, kwargs for the processor)
MODES = {
    "greedy": dict(temperature=1.0, top_k=0, top_p=1.0, greedy=True),
    "multinomial": dict(temperature=1.0, top_k=0, top_p=1.0, greedy=False),
    "temp0.7": dict(temperature=0.7, top_k=0, top_p=1.0, greedy=False),
    "topk50": dict(temperature=0.7, top_k=50, top_p=1.0, greedy=False),
    "nucleus0.9": dict(temperature=1.0, top_k=0, top_p=0.9, greedy=False),
    "nucleus0.9_t0.7": dict(temperature=0.7, top_k=0, top_p=0.9, greedy=False),
    "topk50_p0.9": dict(temperature=0.7, top_k=50, top_p=0.9, greedy=False),
}


def make_case(B, gamma, V, dtype="f32", sigma=0.5, seed=0, scale=3.0, kind="randn", oracle=None, mode=None):
    """-> dict(target [B,g+1,V], draft [B,g,V] in `dtype`, draft_tokens [B,g], u_accept [B,g], u_sample [B]).
    Draft tokens are drawn from the processed drafter distribution with the oracle's own sampler, so
    q_tok > 0 as in real use."""
    g = torch.Generator().manual_seed(1234 + seed)
    t = scale * torch.randn(B, gamma + 1, V, generator=g)
    if kind == "peaked":  # LLM-like: a few dominant tokens
        boost = torch.zeros(B, gamma + 1, V)
        idx = torch.randint(V, (B, gamma + 1, 8), generator=g)
        boost.scatter_(2, idx, 6.0 + 4.0 * torch.rand(B, gamma + 1, 8, generator=g))
        t = t * 0.5 + boost
    d = t[:, :gamma] + sigma * torch.randn(B, gamma, V, generator=g)
    dt = DTYPES[dtype]
    t = t.to(dt)
    d = d.to(dt)
    ua = torch.rand(B, gamma, generator=torch.Generator().manual_seed(777 + seed))
    us = torch.rand(B, generator=torch.Generator().manual_seed(778 + seed))
    ud = torch.rand(B * gamma, generator=torch.Generator().manual_seed(4321 + seed))
    if oracle is not None and gamma > 0:
        m = dict(MODES[mode or "multinomial"])
        tok, _ = oracle.sample_rows(d.float().numpy().reshape(B * gamma, V), ud.numpy(), **m)
        toks = torch.from_numpy(tok.reshape(B, gamma))
    else:
        toks = torch.randint(V, (B, gamma), generator=g)
    return dict(target=t, draft=d, draft_tokens=toks, u_accept=ua, u_sample=us)
