#!/usr/bin/env python
"""Exact-match rate of the canonical arithmetic (C oracle == the CUDA path, bit for bit) against the RAW
reference arithmetic (oracle/torch_port.py == utils/logits_processor.py + sampling/speculative_decoding.py:107-171 in
torch fp32, checked bit-for-bit against the imported reference) at the headline vocabulary, SURVEY.md section 7
"Bit-exactness": every mismatch of an accept decision or an emitted token is classified as a boundary case.

  accept decision  u vs p/q:  mismatch is "boundary" iff |u - p/q| <= 1e-5 * max(p/q, u)   (p/q in float64)
  top-p modes      either mismatch is "cut boundary" iff the nucleus cut of a row involved is decided by <= 1e-5 of
                   cumulative mass (float64): the reference's fp32 cumsum keeps / drops that marginal token.
  emitted token    inverse CDF: mismatch is "boundary" iff u * total lies within 1e-5 (absolute, normalised CDF units,
                   float64 cumulative sums of the reference's own final distribution) of the CDF interval of the
                   oracle's token, i.e. a shift of the CDF by <= 1e-5 -- the north_star probability tolerance --
                   explains it.

    python scripts/ref_match_rate.py [--seqs 2560] [--mode multinomial] [--kind randn|llm] [--out profiles/r2_ref_match_rate.json]
CPU only (test infrastructure; nothing here is on the product path)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import MODES  # noqa: E402
from oracle import oracle, torch_port  # noqa: E402


def make_rows(B, g, V, kind, seed, dtype=torch.bfloat16):
    gen = torch.Generator().manual_seed(31000 + seed)
    t = 3.0 * torch.randn(B, g + 1, V, generator=gen)
    if kind == "llm":
        t = t * 0.5
        idx = torch.randint(V, (B, g + 1, 24), generator=gen)
        t.scatter_(2, idx, 12.0 + 8.0 * torch.rand(B, g + 1, 24, generator=gen))
    d = t[:, :g] + 0.5 * torch.randn(B, g, V, generator=gen)
    return t.to(dtype).float(), d.to(dtype).float(), gen


def f64_reference(t_row, d_row, mode):
    """float64 evaluation of the reference's formulas for one position (plain / temperature modes only)."""
    T = mode["temperature"]
    return torch.softmax(t_row.double() / T, -1), torch.softmax(d_row.double() / T, -1)


def cut_margin(row, mode):
    """top-p modes: distance (float64, probability units) between top_p and the nearest cumulative-mass boundary of
    the reference's own rule (softmax of the top-k-filtered, descending-sorted logits at T = 1, cumsum, shifted mask,
    utils/logits_processor.py:73-78): the kept set is decided by less than this margin."""
    z = row.double()
    if mode["top_k"]:
        kth = torch.topk(z, mode["top_k"])[0][-1]
        z = torch.where(z < kth, torch.full_like(z, -1e20), z)
    sl, _ = torch.sort(z, descending=True, stable=True)
    cum = torch.cumsum(torch.softmax(sl, -1), -1)
    return float((cum - mode["top_p"]).abs().min())


def run(n_seqs, mode_name, kind, V=128256, g=4, chunk=32, seed=0, verbose=True):
    mode = MODES[mode_name]
    torch_port.STABLE_SORT = True  # tie order pinned as in oracle/ref_harness.py (unspecified in the reference)
    plain = mode["top_k"] == 0 and mode["top_p"] >= 1.0
    st = dict(decisions=0, accept_match=0, accept_boundary=0, accept_cut_boundary=0, accept_unexplained=0,
              sequences=0, n_match=0, token_match=0, token_boundary=0, token_cut_boundary=0, token_unexplained=0,
              token_compared=0)
    worst_accept, worst_cdf = 0.0, 0.0
    t0 = time.time()
    for c0 in range(0, n_seqs, chunk):
        B = min(chunk, n_seqs - c0)
        t, d, gen = make_rows(B, g, V, kind, seed * 1000 + c0)
        ud = torch.rand(B * g, generator=gen)
        toks, _ = oracle.sample_rows(d.numpy().reshape(B * g, V), ud.numpy(), **mode)
        toks = torch.from_numpy(toks.reshape(B, g))
        ua, us = torch.rand(B, g, generator=gen), torch.rand(B, generator=gen)
        o = oracle.verify(t, d, toks, ua, us, **mode)
        for b in range(B):
            n_ref, x_ref, frac, p_p = torch_port.verify_one(t[b], d[b], toks[b], mode, r=ua[b], u_sample=float(us[b]),
                                                            want_detail=True)
            st["sequences"] += 1
            # per-position accept tests (the reference evaluates them up to its first rejection; compare all
            # positions both sides define: mask is the per-position test)
            acc_ref = ~(ua[b] > frac)
            for i in range(g):
                st["decisions"] += 1
                a_or = bool(o.accept_mask[b, i])
                if a_or == bool(acc_ref[i]):
                    st["accept_match"] += 1
                    continue
                if plain:
                    p64, q64 = f64_reference(t[b, i], d[b, i], mode)
                    r64 = float(p64[toks[b, i]] / q64[toks[b, i]])
                else:
                    r64 = float(frac[i])
                rel = abs(float(ua[b, i]) - r64) / max(r64, float(ua[b, i]), 1e-30)
                if rel > 1e-5 and not plain and min(cut_margin(t[b, i], mode), cut_margin(d[b, i], mode)) <= 1e-5:
                    st["accept_cut_boundary"] += 1  # a marginal token of the nucleus is kept by one side only
                    continue
                worst_accept = max(worst_accept, rel)
                st["accept_boundary" if rel <= 1e-5 else "accept_unexplained"] += 1
            if n_ref != int(o.n_accepted[b]):
                continue
            st["n_match"] += 1
            st["token_compared"] += 1
            x_or = int(o.next_token[b])
            if x_or == x_ref:
                st["token_match"] += 1
                continue
            # CDF-boundary classification in float64 on the reference's own final distribution
            cum = torch.cumsum(p_p.double().clamp_min(0), 0)
            tgt = float(us[b]) * float(cum[-1])
            # smallest shift of the CDF (float64, normalised) that moves u * total into the oracle token's interval
            lo_or = float(cum[x_or - 1]) if x_or > 0 else 0.0
            gap = max(0.0, lo_or - tgt, tgt - float(cum[x_or])) / float(cum[-1])
            if gap > 1e-5 and not plain:
                rows = [t[b, n_ref]] + ([d[b, n_ref]] if n_ref < g else [])
                if min(cut_margin(r_, mode) for r_ in rows) <= 1e-5:
                    st["token_cut_boundary"] += 1
                    continue
            worst_cdf = max(worst_cdf, gap)
            st["token_boundary" if gap <= 1e-5 else "token_unexplained"] += 1
        if verbose:
            print(f"[{mode_name}/{kind}] {st['sequences']}/{n_seqs} sequences, {time.time() - t0:.0f}s", file=sys.stderr)
    st.update(mode=mode_name, kind=kind, V=V, gamma=g, dtype="bf16 values", worst_accept_rel=worst_accept, worst_cdf_shift=worst_cdf,
              accept_match_rate=st["accept_match"] / max(1, st["decisions"]),
              accepted_length_match_rate=st["n_match"] / max(1, st["sequences"]),
              token_match_rate=st["token_match"] / max(1, st["token_compared"]))
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=2560)
    ap.add_argument("--mode", default="multinomial")
    ap.add_argument("--kind", default="both")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_ref_match_rate.json"))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    res = []
    for kind in (("randn", "llm") if a.kind == "both" else (a.kind,)):
        res.append(run(a.seqs, a.mode, kind))
        print(json.dumps(res[-1]))
    if a.out:
        old = json.load(open(a.out)) if os.path.exists(a.out) else []
        old = [r for r in old if not any(r["mode"] == n["mode"] and r["kind"] == n["kind"] for n in res)]
        json.dump(old + res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
