"""torch-eager PORT of the reference's CPU verify arithmetic (TEST / BASELINE INFRASTRUCTURE ONLY).

/root/reference cannot travel to the GPU box, so bench.py's cpu_baseline and `--impl reference` legs
time this restatement instead: the same torch ops, in the same order, as
  utils/logits_processor.py:13-15 (softmax(_process(x)/T)), :35-36 (argmax), :48-49 (multinomial),
  :59-63 (top-k), :73-81 (nucleus: sort, cumsum(softmax), shifted mask, un-sort by argsort+gather),
  :92-103 (both), and sampling/speculative_decoding.py:107,120-124,135-171 (q capture, p, rand,
  p/q, first rejection, max_fn residual, bonus row, sample).
It is checked against the imported reference in tests/test_oracle_vs_reference.py (build container).
Nothing under speculative-decoding_b200/ imports it."""
import torch
from torch.nn import functional as F


# torch.sort's order of EQUAL logits is unspecified (it differs between torch builds / thread counts); the parity
# harness pins it to the stable order, exactly as oracle/ref_harness.py:stable_sort() does for the imported reference.
STABLE_SORT = False


def _process(logits, mode):
    k, p = mode["top_k"], mode["top_p"]
    if k and k > 0:  # utils/logits_processor.py:59-63 (the reference mutates in place; callers clone)
        top_k = min(k, logits.size(-1))
        rm = logits < torch.topk(logits, top_k, dim=-1)[0][..., -1, None]
        logits = logits.clone()
        logits[rm] = -1e20
    if p is not None and 0.0 < p < 1.0:  # utils/logits_processor.py:73-81
        sorted_logits, sorted_indices = torch.sort(logits, descending=True, stable=STABLE_SORT)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        rem = cumulative_probs > p
        rem[..., 1:] = rem[..., :-1].clone()
        rem[..., 0] = 0
        sorted_logits[rem] = -1e20
        logits = torch.gather(sorted_logits, -1, sorted_indices.argsort(-1))
    return logits


def processor(logits, mode):
    return F.softmax(_process(logits, mode) / mode["temperature"], dim=-1)  # utils/logits_processor.py:13-15


def sample(probs, mode, gen=None):
    if mode["greedy"]:
        return torch.argmax(probs, dim=-1).unsqueeze(-1)  # :35-36
    try:
        return torch.multinomial(probs, num_samples=1, generator=gen)  # :48-49
    except RuntimeError:
        return torch.multinomial(probs.float(), num_samples=1, generator=gen)


def max_fn(x):  # sampling/speculative_decoding.py:10-19
    x_max = torch.where(x > 0, x, torch.zeros_like(x))
    return x_max / torch.sum(x_max, dim=-1, keepdim=True)


def sample_rows(draft_logits, mode, gen=None):
    """drafter side (sampling/speculative_decoding.py:120-124) for g rows -> tokens [g]"""
    return torch.stack([sample(processor(draft_logits[k:k + 1], mode), mode, gen).reshape(()) for k in
                        range(draft_logits.shape[0])])


def inv_cdf64(probs, u):
    """sample() restated as inverse CDF on an injected uniform: float64 cumulative sums of the
    reference's fp32 probabilities, first index whose cumulative mass exceeds u * total."""
    p = probs.detach().double().reshape(-1).clamp_min(0)
    cum = torch.cumsum(p, 0)
    j = int(torch.searchsorted(cum, torch.tensor(float(u) * float(cum[-1]), dtype=torch.float64), right=True))
    nz = torch.nonzero(p > 0).reshape(-1)
    return min(j, int(nz[-1])) if nz.numel() else 0


def verify_one(t_logits, d_logits, toks, mode, gen=None, r=None, skip_sample_adjustment=False, u_sample=None,
               want_detail=False):
    """One verify step of one sequence: t_logits [g+1,V], d_logits [g,V], toks [g] -> (n, x).
    u_sample: the final sample() as inverse CDF on this uniform instead of torch.multinomial;
    want_detail: also return (fractions at the draft tokens [g], p_p)."""
    g, V = d_logits.shape
    q = torch.zeros((1, g, V))                                    # :107 (fp32)
    for k in range(g):
        q[0, k] = processor(d_logits[k:k + 1], mode)[0]           # :121-122
    p = processor(t_logits[:g].unsqueeze(0), mode)                # :135-136
    if r is None:
        r = torch.rand(g, generator=gen)                          # :139
    fractions = p / q                                             # :140
    n = g
    for i in range(g):                                            # :142-145
        if r[i] > fractions[0, i, toks[i]]:
            n = i
            break
    if n == g:
        p_p = processor(t_logits[g:g + 1], mode)                  # :158-160
    elif not skip_sample_adjustment:
        p_p = max_fn(p[..., n, :] - q[0, n, :])                   # :168
    else:
        p_p = p[..., n, :]
    if u_sample is not None and not mode["greedy"]:
        x = inv_cdf64(p_p, u_sample)
    else:
        x = int(sample(p_p, mode, gen).reshape(-1)[0])            # :171
    if want_detail:
        return n, x, fractions[0, torch.arange(g), toks].clone(), p_p.reshape(-1)
    return n, x
