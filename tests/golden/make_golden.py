#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Only runnable in the build container (the reference tree does not travel to the GPU box); the
fixtures it writes are committed.  Inputs are NOT stored: they are re-generated from the seeds by
tests/cases.py / the functions below, so the fixtures only hold the reference's outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh  # noqa: E402
from cases import MODES, make_case  # noqa: E402

KIND = {"greedy": "greedy", "multinomial": "multinomial", "temp0.7": "multinomial", "topk50": "topk",
        "nucleus0.9": "nucleus", "nucleus0.9_t0.7": "nucleus", "topk50_p0.9": "topk_nucleus"}


def tables(seed, L, V, sigma=0.7, kind="peaked", dtype=torch.float32, force=None):
    """Position-indexed target / drafter logit tables [L, V] (float32 values, optionally bf16-rounded).
    kind "llm": LLM-like rows (24 dominant tokens carry almost all the mass), the shape real models emit at
    V=128256; force = "pos:token,..." puts a one-hot row at those table positions in BOTH tables."""
    g = torch.Generator().manual_seed(9000 + seed)
    if kind == "llm":
        p = 1.5 * torch.randn(L, V, generator=g)
        idx = torch.randint(V, (L, 24), generator=g)
        p.scatter_(1, idx, 12.0 + 8.0 * torch.rand(L, 24, generator=g))
        sigma = 0.4
    else:
        p = 2.0 * torch.randn(L, V, generator=g)
        if kind == "peaked":
            idx = torch.randint(V, (L, 6), generator=g)
            p.scatter_(1, idx, 5.0 + 3.0 * torch.rand(L, 6, generator=g))
    q = p + sigma * torch.randn(L, V, generator=g)
    if force:
        for item in force.split(","):
            pos, tok = (int(x) for x in item.split(":"))
            p[pos] = -20.0; p[pos, tok] = 20.0
            q[pos] = p[pos]
    return q.to(dtype).float(), p.to(dtype).float()


def parse_eos(v):
    """eos field of a golden key: "-1", "7" or a "+"-separated list "33+44" (order matters, see
    sampling/speculative_decoding.py:150-152)."""
    v = str(v)
    return [int(x) for x in v.split("+")] if "+" in v else int(v)


def gen_processor_cases():
    """LogitsProcessor.__call__ of the reference on seeded rows: summary statistics of the output."""
    mods = rh.ref_modules()
    out = {}
    for V in (1000, 32000, 128256):
        for dtype in ("f32", "bf16"):
            for kind in ("randn", "peaked"):
                z = make_case(B=2, gamma=1, V=V, dtype=dtype, seed=V % 89, kind=kind)["target"].float()
                for mode, m in MODES.items():
                    proc = rh.make_processor(KIND[mode], m["temperature"], m["top_k"], m["top_p"])
                    with rh.stable_sort(mods.lp):
                        pr = proc(z.clone()).reshape(-1, V)
                    key = f"proc|{V}|{dtype}|{kind}|{mode}"
                    gi = torch.Generator().manual_seed(5)
                    probe = torch.randint(V, (256,), generator=gi)
                    top = pr.topk(32, dim=-1)
                    out[key + "|nkept"] = (pr > 0).sum(-1).numpy()
                    out[key + "|top_idx"] = top.indices.numpy()
                    out[key + "|top_val"] = top.values.numpy()
                    out[key + "|probe"] = pr[:, probe].numpy()
                    out[key + "|argmax"] = pr.argmax(-1).numpy()
    np.savez_compressed(os.path.join(HERE, "processors.npz"), **out)
    print("processors.npz:", len(out), "arrays")


def gen_specgen_cases():
    """speculative_generate end to end on fake models (sampling/speculative_decoding.py:23)."""
    out = {}
    cfgs = []
    for i, mode in enumerate(MODES):
        cfgs.append(dict(seed=i, V=997 if i % 2 else 2048, gamma=4 + (i % 3), mode=mode, max_gen_len=40, skip=False,
                         dtype="bf16" if i % 2 else "f32"))
    cfgs.append(dict(seed=20, V=1500, gamma=5, mode="multinomial", max_gen_len=30, skip=True, dtype="f32"))
    cfgs.append(dict(seed=21, V=1500, gamma=3, mode="topk50_p0.9", max_gen_len=25, skip=False, dtype="bf16", eos=7))
    # two stop tokens, the SECOND-listed one accepted earlier in the window than the first-listed one: the reference
    # truncates at the first-listed token (torch.nonzero row order, sampling/speculative_decoding.py:150-152)
    cfgs.append(dict(seed=22, V=1200, gamma=5, mode="multinomial", max_gen_len=20, skip=False, dtype="f32", eos="33+44",
                     force="7:44,9:33", sigma=0.0))
    cfgs.append(dict(seed=23, V=1200, gamma=5, mode="greedy", max_gen_len=20, skip=False, dtype="f32", eos="44+33",
                     force="7:44,9:33", sigma=0.0))
    # the headline vocabulary (Llama-3, V=128256), LLM-like rows, bf16 values
    cfgs.append(dict(seed=24, V=128256, gamma=4, mode="multinomial", max_gen_len=14, skip=False, dtype="bf16", kind="llm"))
    cfgs.append(dict(seed=25, V=128256, gamma=4, mode="nucleus0.9", max_gen_len=14, skip=False, dtype="bf16", kind="llm"))
    cfgs.append(dict(seed=26, V=128256, gamma=4, mode="topk50_p0.9", max_gen_len=14, skip=False, dtype="bf16", kind="llm"))
    for c in cfgs:
        m = MODES[c["mode"]]
        P = 6
        L = P + c["max_gen_len"] + 2
        q, p = tables(c["seed"], L, c["V"], dtype=torch.bfloat16 if c["dtype"] == "bf16" else torch.float32,
                      kind=c.get("kind", "peaked"), force=c.get("force"), sigma=float(c.get("sigma", 0.7)))
        rng = np.random.RandomState(100 + c["seed"])
        su, au = rng.rand(4096).astype(np.float32), rng.rand(4096).astype(np.float32)
        prompt = rng.randint(0, c["V"], size=P).tolist()
        toks, rate, ns, na = rh.run_speculative_generate(
            prompt, q, p, KIND[c["mode"]], gamma=c["gamma"], max_gen_len=c["max_gen_len"], temperature=m["temperature"],
            top_k=m["top_k"], top_p=m["top_p"], eos=parse_eos(c.get("eos", -1)), skip_sample_adjustment=c["skip"], sample_u=su,
            accept_u=au)
        key = "spec|" + "|".join(f"{k}={c[k]}" for k in sorted(c))
        out[key + "|tokens"] = np.asarray(toks, np.int64)
        out[key + "|rate"] = np.asarray([rate], np.float64)
        out[key + "|used"] = np.asarray([ns, na], np.int64)
    np.savez_compressed(os.path.join(HERE, "specgen.npz"), **out)
    print("specgen.npz:", len(out) // 3, "runs")


def gen_ngram_cases():
    out = {}
    for i, mode in enumerate(["greedy", "multinomial", "topk50"]):
        c = dict(seed=30 + i, V=64, gamma=4, mode=mode, max_gen_len=48, n=3 + (i % 2), filler=3 if i != 1 else 1)
        m = MODES[mode]
        P = 12
        L = P + c["max_gen_len"] + 2
        g = torch.Generator().manual_seed(c["seed"])
        # strongly position-periodic logits => repeating text the n-gram drafter can learn
        base = torch.zeros(L, c["V"])
        for t in range(L):
            base[t, (3 * (t % 5)) % c["V"]] = 9.0
        p = base + 0.3 * torch.randn(L, c["V"], generator=g)
        rng = np.random.RandomState(c["seed"])
        su = rng.rand(4096).astype(np.float32)
        fb = rng.randint(0, c["V"], size=512)
        prompt = [(3 * (t % 5)) % c["V"] for t in range(1, P + 1)]
        toks, rate, ns, nfb = rh.run_ngram_generate(prompt, p, KIND[mode], ngram_n=c["n"], gamma=c["gamma"],
                                                    max_gen_len=c["max_gen_len"], filler_top_k=c["filler"],
                                                    temperature=m["temperature"], top_k=m["top_k"], top_p=m["top_p"],
                                                    sample_u=su, fallback_tokens=fb)
        key = "ngram|" + "|".join(f"{k}={c[k]}" for k in sorted(c))
        out[key + "|tokens"] = np.asarray(toks, np.int64)
        out[key + "|rate"] = np.asarray([rate], np.float64)
        out[key + "|used"] = np.asarray([ns, nfb], np.int64)
    # NGramStorage / OneLevelNGramStorage driven directly
    mods = rh.ref_modules()
    for one in (0, 1):
        rng = np.random.RandomState(7 + one)
        cls = mods.ng.OneLevelNGramStorage if one else mods.ng.NGramStorage
        st = cls(4, 50)
        seqs = rng.randint(0, 5, size=(3, 24))
        st.initialize(torch.from_numpy(seqs))
        res = []
        for step in range(10):
            nt = rng.randint(0, 5, size=(3, 1 + step % 3))
            st.update(torch.from_numpy(seqs[:, :24 - step]), torch.from_numpy(nt))
            import ngram_assisted.ngram_storage as ngs
            with rh.patched(ngs, randint=lambda high, size=None, **k: torch.full(size, 49, dtype=torch.long)):
                tok, known = st.next_token(torch.from_numpy(seqs[:, :24 - step]))
            res.append(np.concatenate([tok.numpy(), known.numpy().astype(np.int64)]))
        out[f"ngramtable|one={one}"] = np.stack(res)
    np.savez_compressed(os.path.join(HERE, "ngram.npz"), **out)
    print("ngram.npz:", len(out), "arrays")


def gen_batch_cases():
    out = {}
    for i in range(3):
        c = dict(seed=40 + i, V=512, gamma=3 + i, B=3, gen_len=20, P=5, end=11 if i == 2 else -1)
        g = torch.Generator().manual_seed(c["seed"])
        L = c["P"] + c["gen_len"] + 2
        p = 2.0 * torch.randn(c["B"], L, c["V"], generator=g)
        idx = torch.randint(c["V"], (c["B"], L, 4), generator=g)
        p.scatter_(2, idx, 6.0 + 2.0 * torch.rand(c["B"], L, 4, generator=g))
        q = p + 0.7 * torch.randn(c["B"], L, c["V"], generator=g)
        rng = np.random.RandomState(c["seed"])
        su, au = rng.rand(8192).astype(np.float32), rng.rand(8192).astype(np.float32)
        ids = torch.from_numpy(rng.randint(1, c["V"], size=(c["B"], c["P"])))
        outs, rates, ns, na = rh.run_batch_speculative_generate(ids, q, p, gamma=c["gamma"], gen_len=c["gen_len"],
                                                                end_tokens=[c["end"]] if c["end"] >= 0 else [],
                                                                sample_u=su, accept_u=au)
        key = "batch|" + "|".join(f"{k}={c[k]}" for k in sorted(c))
        for b, o in enumerate(outs):
            out[key + f"|out{b}"] = np.asarray(o, np.int64)
        out[key + "|rates"] = np.asarray(rates, np.float64)
        out[key + "|used"] = np.asarray([ns, na], np.int64)
    np.savez_compressed(os.path.join(HERE, "batch.npz"), **out)
    print("batch.npz:", len(out), "arrays")


if __name__ == "__main__":
    assert rh.available(), "reference tree not found"
    torch.manual_seed(0)
    gen_processor_cases()
    gen_specgen_cases()
    gen_ngram_cases()
    gen_batch_cases()
