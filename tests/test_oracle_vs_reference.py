"""CPU, build container only: live cross-checks against the imported reference (/root/reference).
Skipped on the GPU box, where the reference tree does not exist."""
import os
import sys

import numpy as np
import pytest
import torch

from cases import MODES, make_case

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")

KIND = {"greedy": "greedy", "multinomial": "multinomial", "temp0.7": "multinomial", "topk50": "topk",
        "nucleus0.9": "nucleus", "nucleus0.9_t0.7": "nucleus", "topk50_p0.9": "topk_nucleus"}


@pytest.mark.parametrize("mode", list(MODES))
def test_torch_port_equals_reference_processors(mode):
    """oracle/torch_port.py (the CPU baseline bench.py times) is the reference's arithmetic, bit for bit."""
    from oracle import torch_port
    mods = rh.ref_modules()
    m = MODES[mode]
    z = make_case(B=2, gamma=2, V=5000, dtype="f32", seed=2)["target"].float()
    ref = rh.make_processor(KIND[mode], m["temperature"], m["top_k"], m["top_p"])(z.clone())
    assert torch.equal(torch_port.processor(z.clone(), m), ref)
    x = torch.randn(3, 50)
    assert torch.equal(torch_port.max_fn(x), mods.sd.max_fn(x))


def test_torch_port_verify_matches_reference_loop():
    """one speculative step of the reference == torch_port.verify_one on the same tables / uniforms."""
    from oracle import torch_port
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    V, gamma = 600, 4
    q, p = make_golden.tables(77, 20, V)
    rng = np.random.RandomState(1)
    su, au = rng.rand(256).astype(np.float32), rng.rand(256).astype(np.float32)
    prompt = [5, 6, 7]
    toks, rate, ns, na = rh.run_speculative_generate(prompt, q, p, "multinomial", gamma=gamma, max_gen_len=6,
                                                     first_target=False, sample_u=su, accept_u=au)
    # replay the first step with the port: drafts are the first n tokens (+ the corrected one)
    mode = MODES["multinomial"]
    drafts = []
    si = 0
    for k in range(gamma):
        pr = torch_port.processor(q[2 + k:3 + k], mode)
        drafts.append(rh.inv_cdf_reference(pr[0], su[si])); si += 1
    n, _ = torch_port.verify_one(p[2:2 + gamma + 1], q[2:2 + gamma], torch.tensor(drafts), mode,
                                 r=torch.from_numpy(au[:gamma].copy()))
    assert toks[:n] == drafts[:n]


def test_oracle_accept_decisions_match_reference_arithmetic(oracle_mod):
    """accept decisions and lengths of the canonical arithmetic == the reference's fp32 torch arithmetic
    on seeded cases (differences could only arise for |u - p/q| < ~1e-5 p/q)."""
    from oracle import torch_port
    for mode in ("multinomial", "topk50", "topk50_p0.9"):
        m = MODES[mode]
        case = make_case(B=16, gamma=4, V=4096, dtype="f32", sigma=0.6, seed=12, oracle=oracle_mod, mode=mode, kind="peaked")
        o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
        for b in range(16):
            n, _ = torch_port.verify_one(case["target"][b], case["draft"][b], case["draft_tokens"][b], m,
                                         r=case["u_accept"][b])
            assert n == o.n_accepted[b], (mode, b)


@pytest.mark.parametrize("one_level", [False, True])
def test_ngram_oracle_has_gram_matches_reference_classes(one_level):
    """ngram_storage.py:98-106 / :181-193 on the real classes vs oracle/ngram_oracle.py:has_gram."""
    from oracle.ngram_oracle import NGramOracle
    ng = rh.ref_modules().ng
    import ngram_assisted.ngram_storage as ns
    rng = np.random.RandomState(3)
    n, V = 4, 50
    ref = (ns.OneLevelNGramStorage if one_level else ns.NGramStorage)(n, V)
    orc = NGramOracle(n, V, one_level)
    seq = rng.randint(0, 5, size=(1, 40))
    ref.initialize(torch.from_numpy(seq))
    orc.initialize([seq[0]])
    for step in range(6):
        nt = rng.randint(0, 5, size=(1, 2))
        ref.update(torch.from_numpy(seq[:, :30 + step]), torch.from_numpy(nt))
        orc.update([seq[0, :30 + step]], nt)
    for trial in range(300):
        q = rng.randint(0, 5, size=int(rng.randint(1, 8)))
        try:
            want = bool(ref.has_gram(torch.from_numpy(q)))
        except KeyError:  # the reference indexes counts[j] of a level it never created (a miss)
            want = False
        assert orc.has_gram(q) == want, q
