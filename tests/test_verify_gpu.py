"""Parity of the CUDA verify path (through the C ABI / torch ops) with the CPU oracle: accepted
lengths, emitted tokens, accept masks, stop indices BIT-EXACT; probabilities within 1e-5 (they are in
fact bit-identical because both sides use the same canonical arithmetic)."""
import numpy as np
import pytest
import torch

from cases import MODES, make_case

pytestmark = pytest.mark.gpu

F_BATCHED, F_NO_BONUS, F_SKIP, F_NGRAM, F_FALLBACK = 1, 2, 4, 8, 16
TMA_NGRAM_DEFAULT = 1  # library default of the "tma_ngram" option (tests that flip it restore this)


def _run_both(oracle, case, mode, flags=0, stop=(), dev="cuda"):
    import specdec_b200 as sd
    m = MODES[mode]
    ngram = bool(flags & F_NGRAM)
    tgt = case["target"][:, :-1] if (flags & F_NO_BONUS) else case["target"]
    o = oracle.verify(tgt if not (flags & F_NO_BONUS) else case["target"], None if ngram else case["draft"],
                      case["draft_tokens"], case["u_accept"], case["u_sample"], flags=flags, stop_tokens=stop, **m)
    r = sd.fused_verify(tgt.to(dev), None if ngram else case["draft"].to(dev), case["draft_tokens"].to(dev),
                        case["u_accept"].to(dev), case["u_sample"].to(dev), flags=flags, stop_tokens=list(stop), **m)
    torch.cuda.synchronize()
    if tgt.shape[0] <= 64:
        # every small case is ALSO run through the opt-in one-launch cluster path (csrc/cluster_small.cuh; 16- and 8-CTA
        # clusters) and through the pipeline with the atomics + counters tail: the decisions must be identical
        lib = sd._lib.lib()
        for opts in ({b"small_b": 64}, {b"small_b": 64, b"small_cl": 8}, {b"tail_slots": 0}, {b"static_rows": 1}):
            opt = tuple(opts.items())
            for k_, v_ in opts.items():
                assert lib.specdec_set_option(k_, v_) == 0
            r2 = sd.fused_verify(tgt.to(dev), None if ngram else case["draft"].to(dev), case["draft_tokens"].to(dev),
                                 case["u_accept"].to(dev), case["u_sample"].to(dev), flags=flags, stop_tokens=list(stop), **m)
            torch.cuda.synchronize()
            assert lib.specdec_set_option(b"small_b", 0) == 0 and lib.specdec_set_option(b"small_cl", 16) == 0
            assert lib.specdec_set_option(b"tail_slots", 1) == 0 and lib.specdec_set_option(b"static_rows", 0) == 0
            for a in ("n_accepted", "next_token", "accept_mask", "first_stop", "packed"):
                assert torch.equal(getattr(r, a), getattr(r2, a)), (a, opt)
            np.testing.assert_allclose(r2.p_tok.cpu().numpy(), r.p_tok.cpu().numpy(), rtol=1e-5, atol=0)
    return o, r


def _assert_same(o, r, ngram=False):
    assert np.array_equal(r.n_accepted.cpu().numpy(), o.n_accepted), "accepted lengths differ"
    assert np.array_equal(r.next_token.cpu().numpy(), o.next_token), "emitted tokens differ"
    assert np.array_equal(r.accept_mask.cpu().numpy(), o.accept_mask), "accept masks differ"
    assert np.array_equal(r.first_stop.cpu().numpy(), o.first_stop), "stop index differs"
    np.testing.assert_allclose(r.p_tok.cpu().numpy(), o.p_tok, rtol=1e-5, atol=0)
    if not ngram:
        np.testing.assert_allclose(r.q_tok.cpu().numpy(), o.q_tok, rtol=1e-5, atol=0)
    # packed = {n, accepted drafts, next, -1...}
    pk = r.packed.cpu().numpy()
    assert np.array_equal(pk[:, 0], o.n_accepted)
    for b in range(pk.shape[0]):
        n = int(o.n_accepted[b])
        assert pk[b, 1 + n] == o.next_token[b]


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("V", [1000, 32000])
def test_verify_modes(oracle_mod, mode, dtype, V):
    case = make_case(B=6, gamma=4, V=V, dtype=dtype, sigma=0.5, seed=V % 97, oracle=oracle_mod, mode=mode)
    o, r = _run_both(oracle_mod, case, mode)
    _assert_same(o, r)


@pytest.mark.parametrize("mode", ["multinomial", "topk50", "nucleus0.9", "topk50_p0.9", "greedy"])
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_verify_llama_vocab(oracle_mod, mode, dtype):
    case = make_case(B=8, gamma=4, V=128256, dtype=dtype, sigma=0.5, seed=5, oracle=oracle_mod, mode=mode)
    o, r = _run_both(oracle_mod, case, mode)
    _assert_same(o, r)
    assert 0 < o.n_accepted.sum() < 8 * 4  # the case exercises both accept and reject


@pytest.mark.parametrize("V", [7, 257, 4099, 50257])  # ragged: odd V => rows not 16-byte aligned
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("mode", ["multinomial", "topk50_p0.9", "nucleus0.9"])
def test_verify_ragged_vocab(oracle_mod, V, dtype, mode):
    case = make_case(B=3, gamma=3, V=V, dtype=dtype, sigma=1.0, seed=V, oracle=oracle_mod, mode=mode)
    o, r = _run_both(oracle_mod, case, mode)
    _assert_same(o, r)


@pytest.mark.parametrize("gamma", [1, 2, 5, 8])
@pytest.mark.parametrize("sigma", [0.0, 3.0])
def test_verify_gamma_and_acceptance_extremes(oracle_mod, gamma, sigma):
    case = make_case(B=5, gamma=gamma, V=8192, dtype="bf16", sigma=sigma, seed=gamma, oracle=oracle_mod)
    o, r = _run_both(oracle_mod, case, "multinomial")
    _assert_same(o, r)
    if sigma == 0.0:
        assert (o.n_accepted == gamma).all()  # p == q: everything accepted, bonus row drawn


@pytest.mark.parametrize("flags", [F_SKIP, F_BATCHED | F_NO_BONUS | F_FALLBACK, F_BATCHED, F_NO_BONUS])
@pytest.mark.parametrize("mode", ["multinomial", "greedy", "topk50"])
def test_verify_variant_flags(oracle_mod, flags, mode):
    case = make_case(B=6, gamma=4, V=32000, dtype="f32", sigma=0.7, seed=11 + flags, oracle=oracle_mod, mode=mode)
    o, r = _run_both(oracle_mod, case, mode, flags=flags)
    _assert_same(o, r)


@pytest.mark.parametrize("mode", ["greedy", "multinomial", "topk50"])
def test_verify_ngram_mode(oracle_mod, mode):
    case = make_case(B=6, gamma=4, V=32000, dtype="bf16", sigma=0.0, seed=3, kind="peaked")
    m = MODES[mode]
    # drafts = what the target itself would pick for the first positions, then garbage => partial accepts
    tok, _ = oracle_mod.sample_rows(case["target"][:, :4].float().numpy().reshape(24, 32000),
                                    case["u_accept"].numpy().reshape(-1), **m)
    toks = torch.from_numpy(tok.reshape(6, 4)).clone()
    toks[::2, 2] = 17
    case["draft_tokens"] = toks
    o, r = _run_both(oracle_mod, case, mode, flags=F_NGRAM)
    _assert_same(o, r, ngram=True)
    assert o.n_accepted.max() == 4 and o.n_accepted.min() == 2


def test_verify_stop_tokens(oracle_mod):
    case = make_case(B=8, gamma=4, V=4096, dtype="f32", sigma=0.0, seed=21, oracle=oracle_mod)
    stop = [int(case["draft_tokens"][0, 1]), int(case["draft_tokens"][3, 0]), 4095]
    o, r = _run_both(oracle_mod, case, "multinomial", stop=stop)
    _assert_same(o, r)
    assert o.first_stop[0] in (0, 1) and o.first_stop[3] == 0


def test_verify_adversarial_rows(oracle_mod):
    """all-equal logits (ties everywhere), one-hot rows, -inf entries, p == q."""
    B, g, V = 4, 3, 2048
    t = torch.zeros(B, g + 1, V)
    d = torch.zeros(B, g, V)
    t[1] = -30.0; t[1, :, 5] = 10.0          # one-hot-ish
    d[1] = -30.0; d[1, :, 5] = 10.0
    t[2] = torch.randn(g + 1, V); t[2, :, ::3] = float("-inf")
    d[2] = torch.randn(g, V); d[2, :, ::3] = float("-inf")
    t[3] = torch.randn(g + 1, V).round()     # many exact ties
    d[3] = t[3, :g]
    toks = torch.tensor([[0, 7, 2047], [5, 5, 5], [1, 2, 4], [9, 10, 11]])
    case = dict(target=t, draft=d, draft_tokens=toks, u_accept=torch.tensor([[0.1, 0.5, 0.999]] * B),
                u_sample=torch.tensor([0.0, 0.3, 0.77, 0.99999]))
    for mode in ["multinomial", "greedy", "topk50", "nucleus0.9", "topk50_p0.9"]:
        o, r = _run_both(oracle_mod, case, mode)
        _assert_same(o, r)


def test_strided_logits_views(oracle_mod):
    """target_logits as a slice of a longer [B, L, V] model output (non-contiguous batch stride)."""
    import specdec_b200 as sd
    case = make_case(B=4, gamma=3, V=4096, dtype="bf16", sigma=0.5, seed=9, oracle=oracle_mod)
    full = torch.zeros(4, 10, 4096, dtype=torch.bfloat16)
    full[:, 5:9] = case["target"]
    view = full.cuda()[:, 5:9]
    m = MODES["multinomial"]
    r = sd.fused_verify(view, case["draft"].cuda(), case["draft_tokens"].cuda(), case["u_accept"].cuda(),
                        case["u_sample"].cuda(), **m)
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
    _assert_same(o, r)


def test_philox_matches_oracle_and_is_sharding_independent(oracle_mod):
    import specdec_b200 as sd
    ua, us = sd.philox_uniform(2025, 7, 100, 16, 4)
    oa, os_ = oracle_mod.philox_uniform(2025, 7, 100, 16, 4)
    assert np.array_equal(ua.cpu().numpy(), oa) and np.array_equal(us.cpu().numpy(), os_)
    ua2, us2 = sd.philox_uniform(2025, 7, 108, 8, 4)  # second half as its own shard
    assert torch.equal(ua2, ua[8:]) and torch.equal(us2, us[8:])
    assert 0.0 <= float(ua.min()) and float(ua.max()) < 1.0
    # in-kernel uniforms == dumped uniforms
    case = make_case(B=16, gamma=4, V=4096, dtype="f32", sigma=0.5, seed=2, oracle=oracle_mod)
    r1 = sd.fused_verify(case["target"].cuda(), case["draft"].cuda(), case["draft_tokens"].cuda(), None, None,
                         seed=2025, offset=7, seq_id0=100)
    r2 = sd.fused_verify(case["target"].cuda(), case["draft"].cuda(), case["draft_tokens"].cuda(), ua, us)
    assert torch.equal(r1.n_accepted, r2.n_accepted) and torch.equal(r1.next_token, r2.next_token)


def test_full_size_properties():
    """BASELINE.json full size (B=256, gamma=4, V=128256, bf16): size-independent properties.
    p == q  =>  every draft accepted for any u; greedy bonus token == argmax of the bonus row;
    shuffling the batch permutes the outputs (results do not depend on the row's CTA)."""
    import specdec_b200 as sd
    B, g, V = 256, 4, 128256
    gen = torch.Generator(device="cuda").manual_seed(0)
    t = (3 * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
    toks = torch.randint(V, (B, g), device="cuda", generator=gen)
    ua = torch.rand(B, g, device="cuda", generator=gen)
    us = torch.rand(B, device="cuda", generator=gen)
    r = sd.fused_verify(t, t[:, :g].contiguous(), toks, ua, us, greedy=True)
    assert bool((r.n_accepted == g).all())
    assert torch.equal(r.next_token, t[:, g].float().argmax(-1))
    np.testing.assert_allclose(r.p_tok.cpu().numpy(), r.q_tok.cpu().numpy(), rtol=0, atol=0)
    ref_p = torch.softmax(t[:, :g].float(), -1).gather(-1, toks.unsqueeze(-1)).squeeze(-1)
    np.testing.assert_allclose(r.p_tok.cpu().numpy(), ref_p.cpu().numpy(), rtol=2e-5, atol=1e-30)
    d = (t[:, :g].float() + 0.5 * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
    r1 = sd.fused_verify(t, d, toks, ua, us)
    perm = torch.randperm(B, device="cuda", generator=gen)
    r2 = sd.fused_verify(t[perm].contiguous(), d[perm].contiguous(), toks[perm], ua[perm], us[perm])
    assert torch.equal(r1.n_accepted[perm], r2.n_accepted) and torch.equal(r1.next_token[perm], r2.next_token)
    assert 0 < int(r1.n_accepted.sum()) < B * g


@pytest.mark.parametrize("mode_kw", [dict(temperature=1.0, top_k=0, top_p=0.9, greedy=False),
                                     dict(temperature=0.8, top_k=2000, top_p=1.0, greedy=False),
                                     dict(temperature=0.8, top_k=3000, top_p=0.95, greedy=False)])
def test_slow_selection_path(oracle_mod, mode_kw):
    """flat / tied rows whose kept set exceeds the shared-memory candidate capacity (8192) or whose
    top_k exceeds 1024 take the sweep-based exact selection."""
    import specdec_b200 as sd
    B, g, V = 2, 1, 20000
    gen = torch.Generator().manual_seed(5)
    t = 0.01 * torch.randn(B, g + 1, V, generator=gen)
    t[1] = 0.0  # every logit tied: nucleus keeps the first ceil(0.9 V) indices
    d = t[:, :g] + 0.01 * torch.randn(B, g, V, generator=gen)
    toks = torch.tensor([[3], [19999]])
    ua = torch.tensor([[0.5], [0.99]]); us = torch.tensor([0.25, 0.9])
    o = oracle_mod.verify(t, d, toks, ua, us, **mode_kw)
    r = sd.fused_verify(t.cuda(), d.cuda(), toks.cuda(), ua.cuda(), us.cuda(), **mode_kw)
    _assert_same(o, r)


def test_tma_and_ldg_row_kernels_agree(oracle_mod):
    """the TMA bulk-copy pipeline and the vectorised-LDG row kernel give identical decisions."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    case = make_case(B=40, gamma=4, V=32000, dtype="bf16", sigma=0.5, seed=31, oracle=oracle_mod)
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
    r1 = sd.fused_verify(*args)
    assert lib.specdec_set_option(b"force_ldg", 1) == 0
    try:
        r2 = sd.fused_verify(*args)
    finally:
        lib.specdec_set_option(b"force_ldg", 0)
    assert torch.equal(r1.n_accepted, r2.n_accepted) and torch.equal(r1.next_token, r2.next_token)
    np.testing.assert_allclose(r1.p_tok.cpu().numpy(), r2.p_tok.cpu().numpy(), rtol=1e-5)
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"])
    _assert_same(o, r1)


@pytest.mark.parametrize("B,gamma,V", [(160, 3, 8192), (200, 4, 32000)])
def test_chunk_pipelining_does_not_change_results(oracle_mod, B, gamma, V):
    """bf16 batches of >= 128 sequences run as two chunks on two streams by default (the path bench.py times):
    "chunks"=1 (single stream), the default (2) and 3 chunks give identical outputs, equal to the oracle."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    case = make_case(B=B, gamma=gamma, V=V, dtype="bf16", sigma=0.6, seed=41, oracle=oracle_mod)
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
    res = {}
    for chunks in (0, 1, 3):  # 0 = library default
        assert lib.specdec_set_option(b"chunks", chunks) == 0
        res[chunks] = sd.fused_verify(*args, stop_tokens=[5, 77])
        torch.cuda.synchronize()
    assert lib.specdec_set_option(b"chunks", 0) == 0
    for other in (1, 3):
        r1, r2 = res[0], res[other]
        for a, b in ((r1.n_accepted, r2.n_accepted), (r1.next_token, r2.next_token), (r1.accept_mask, r2.accept_mask),
                     (r1.first_stop, r2.first_stop), (r1.packed, r2.packed), (r1.p_tok, r2.p_tok), (r1.next_prob, r2.next_prob)):
            assert torch.equal(a, b)
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"],
                          stop_tokens=[5, 77])
    _assert_same(o, res[0])


def _full_size_case(B, gamma, V, mode, seed, oracle_mod, sigma=0.5, kind="randn"):
    """bench.py-shaped inputs built on the CPU in slabs (the oracle needs them on the host anyway)."""
    m = MODES[mode]
    g = torch.Generator().manual_seed(9000 + seed)
    t = torch.empty(B, gamma + 1, V, dtype=torch.bfloat16)
    d = torch.empty(B, gamma, V, dtype=torch.bfloat16)
    for b0 in range(0, B, 32):
        b1 = min(B, b0 + 32)
        tf = 3.0 * torch.randn(b1 - b0, gamma + 1, V, generator=g)
        if kind == "peaked":
            idx = torch.randint(V, (b1 - b0, gamma + 1, 24), generator=g)
            tf = tf * 0.5
            tf.scatter_(2, idx, 12.0 + 8.0 * torch.rand(b1 - b0, gamma + 1, 24, generator=g))
        t[b0:b1] = tf.to(torch.bfloat16)
        d[b0:b1] = (tf[:, :gamma] + sigma * torch.randn(b1 - b0, gamma, V, generator=g)).to(torch.bfloat16)
    ud = torch.rand(B * gamma, generator=g)
    tok, _ = oracle_mod.sample_rows(d.float().numpy().reshape(B * gamma, V), ud.numpy(), **m)
    return dict(target=t, draft=d, draft_tokens=torch.from_numpy(tok.reshape(B, gamma)),
                u_accept=torch.rand(B, gamma, generator=g), u_sample=torch.rand(B, generator=g))


@pytest.mark.parametrize("mode", ["multinomial", "greedy", "topk50", "nucleus0.9"])
def test_headline_shape_matches_oracle(oracle_mod, mode):
    """BASELINE.json headline shape -- B=256, gamma=4, V=128256, bf16, DEFAULT library options (the exact path
    bench.py times) -- against the oracle: accepted lengths, emitted tokens, masks bit-exact."""
    case = _full_size_case(256, 4, 128256, mode, 1, oracle_mod)
    o, r = _run_both(oracle_mod, case, mode)
    _assert_same(o, r)
    assert 0 < int(o.n_accepted.sum()) < 256 * 4


def test_headline_shape_philox_and_repeat_calls(oracle_mod):
    """same shape, uniforms from the in-kernel Philox stream (as bench.py runs it), several back-to-back calls on
    the same stream (workspace reuse across calls), compared with the oracle fed the dumped uniforms."""
    import specdec_b200 as sd
    B, g, V = 256, 4, 128256
    case = _full_size_case(B, g, V, "multinomial", 2, oracle_mod)
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens")]
    rs = [sd.fused_verify(*args, None, None, seed=2025, offset=i, seq_id0=512) for i in range(4)]
    torch.cuda.synchronize()
    for i in (0, 3):
        ua, us = sd.philox_uniform(2025, i, 512, B, g)
        o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], ua.cpu(), us.cpu())
        _assert_same(o, rs[i])


def test_config2_shape_matches_oracle(oracle_mod):
    """BASELINE.json configs[2] shape: batch 64, gamma=4, nucleus p=0.9, V=128256 bf16; LLM-like (peaked) rows and
    the flat rows random-init models emit."""
    for kind, seed in (("peaked", 3), ("randn", 4)):
        case = _full_size_case(64, 4, 128256, "nucleus0.9", seed, oracle_mod, kind=kind)
        o, r = _run_both(oracle_mod, case, "nucleus0.9")
        _assert_same(o, r)


@pytest.mark.parametrize("mode", ["greedy", "multinomial"])
def test_config3_shape_ngram_matches_oracle(oracle_mod, mode):
    """BASELINE.json configs[3] shape: n-gram-assisted verify, batch 128, gamma=6, V=128256 bf16."""
    B, g, V = 128, 6, 128256
    case = _full_size_case(B, g, V, mode, 5, oracle_mod, sigma=0.0, kind="peaked")
    m = MODES[mode]
    tok, _ = oracle_mod.sample_rows(case["target"][:, :g].float().numpy().reshape(B * g, V),
                                    case["u_accept"].numpy().reshape(-1), **m)
    toks = torch.from_numpy(tok.reshape(B, g)).clone()
    toks[::2, 3] = 17
    toks[1::5, 0] = 9
    case["draft_tokens"] = toks
    o, r = _run_both(oracle_mod, case, mode, flags=F_NGRAM)
    _assert_same(o, r, ngram=True)
    assert o.n_accepted.max() == g and o.n_accepted.min() == 0


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
@pytest.mark.parametrize("mode", ["nucleus0.9", "nucleus0.9_t0.7"])
def test_fast_nucleus_path_on_peaked_rows(oracle_mod, dtype, mode):
    """LLM-like rows (small nucleus) are resolved by nucleus_fast_kernel (MUFU normaliser + bracketed
    threshold); results equal the oracle and the exact-only path."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    case = make_case(B=12, gamma=4, V=128256, dtype=dtype, sigma=0.0, seed=17)
    # LLM-like rows: a handful of tokens carry almost all the mass
    g = torch.Generator().manual_seed(5)
    t = case["target"].float() * 0.5
    idx = torch.randint(128256, (12, 5, 24), generator=g)
    t.scatter_(2, idx, 12.0 + 8.0 * torch.rand(12, 5, 24, generator=g))
    d = t[:, :4] + 0.4 * torch.randn(12, 4, 128256, generator=g)
    case["target"], case["draft"] = t.to(case["target"].dtype), d.to(case["target"].dtype)
    m = MODES[mode]
    tok, _ = oracle_mod.sample_rows(case["draft"].float().numpy().reshape(48, 128256),
                                    torch.rand(48, generator=g).numpy(), **m)
    case["draft_tokens"] = torch.from_numpy(tok.reshape(12, 4))
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
    r1 = sd.fused_verify(*args, **m)
    probs_fast, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
    assert lib.specdec_set_option(b"no_fast_nucleus", 1) == 0
    try:
        r2 = sd.fused_verify(*args, **m)
        probs_exact, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
    finally:
        lib.specdec_set_option(b"no_fast_nucleus", 0)
    assert torch.equal(r1.n_accepted, r2.n_accepted) and torch.equal(r1.next_token, r2.next_token)
    assert torch.equal(r1.p_tok, r2.p_tok) and torch.equal(probs_fast, probs_exact)
    kept = (probs_fast > 0).sum(-1)
    assert int(kept.max()) < 1000  # genuinely small nuclei: this case exercises the fast kernel
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
    _assert_same(o, r1)


def test_randomised_configurations(oracle_mod):
    """60 random (B, gamma, V, dtype, processor, flags, sigma) configurations against the oracle."""
    rng = np.random.RandomState(2024)
    modes = list(MODES)
    flag_sets = [0, 0, 0, F_SKIP, F_BATCHED | F_NO_BONUS | F_FALLBACK, F_NO_BONUS, F_BATCHED]
    for it in range(60):
        B = int(rng.choice([1, 2, 3, 5, 9, 17, 33, 70]))
        gamma = int(rng.choice([1, 2, 3, 4, 6, 8]))
        V = int(rng.choice([8, 100, 1001, 4096, 16384, 32000, 50257]))
        dtype = str(rng.choice(["f32", "bf16", "f16"]))
        mode = modes[int(rng.randint(len(modes)))]
        flags = flag_sets[int(rng.randint(len(flag_sets)))]
        sigma = float(rng.choice([0.0, 0.3, 1.0, 3.0]))
        kind = str(rng.choice(["randn", "peaked"]))
        if B * gamma * V > 6e6:
            B = max(1, int(6e6 // (gamma * V)))
        case = make_case(B=B, gamma=gamma, V=V, dtype=dtype, sigma=sigma, seed=1000 + it, kind=kind, oracle=oracle_mod,
                         mode=mode)
        stop = [int(case["draft_tokens"][0, 0])] if it % 3 == 0 else []
        o, r = _run_both(oracle_mod, case, mode, flags=flags, stop=stop)
        try:
            _assert_same(o, r)
        except AssertionError as e:
            raise AssertionError(f"config {it}: B={B} gamma={gamma} V={V} {dtype} {mode} flags={flags} sigma={sigma} {kind}: {e}")


def test_ngram_greedy_fast_path_and_near_ties(oracle_mod):
    """greedy n-gram verify takes the arg-max from the fast row kernel; rows whose two largest logits are
    within the polynomial's resolution fall back to the exact arg-max.  Both equal the oracle."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    B, g, V = 6, 3, 50000
    gen = torch.Generator().manual_seed(8)
    t = 2.0 * torch.randn(B, g + 1, V, generator=gen)
    # near ties: runner-up one ulp (fp32) below / equal to the maximum, at a LOWER and a HIGHER index
    for b in range(B):
        for i in range(g + 1):
            mx = float(t[b, i].max())
            j = int(t[b, i].argmax())
            k = (j + 1234 * (b + 1)) % V if (b + i) % 2 else (j - 777 * (i + 1)) % V
            t[b, i, k] = float(np.nextafter(np.float32(mx), np.float32(-1e9))) if b % 3 else mx
    for dtype in (torch.float32, torch.bfloat16):
        tt = t.to(dtype)
        toks = tt[:, :g].float().argmax(-1)
        toks[1, 1] = 3
        ua = torch.zeros(B, g); us = torch.zeros(B)
        o = oracle_mod.verify(tt, None, toks, ua, us, greedy=True, flags=F_NGRAM)
        r1 = sd.fused_verify(tt.cuda(), None, toks.cuda(), ua.cuda(), us.cuda(), greedy=True, flags=F_NGRAM)
        assert lib.specdec_set_option(b"no_fast_ngram", 1) == 0
        try:
            r2 = sd.fused_verify(tt.cuda(), None, toks.cuda(), ua.cuda(), us.cuda(), greedy=True, flags=F_NGRAM)
        finally:
            lib.specdec_set_option(b"no_fast_ngram", 0)
        _assert_same(o, r1, ngram=True)
        _assert_same(o, r2, ngram=True)
        assert torch.equal(r1.next_token, r2.next_token)


@pytest.mark.parametrize("dtype,V,greedy", [("bf16", 32000, False), ("f32", 50257, False), ("bf16", 128256, True),
                                            ("f16", 4099, False), ("bf16", 151936, False)])
def test_fused_tail_equals_split_kernels_incl_ambiguous_accept_tests(oracle_mod, dtype, V, greedy):
    """tail_fused_kernel (weights of the deciding row pair cached in shared memory, 16 CTAs per sequence) gives
    the same outputs as exact_rows + sample_partial and the oracle -- including sequences whose accept test
    falls inside the fast path's safety margin (u within 1e-4 of p/q), which take the in-kernel exact route."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    B, g = 24, 4
    case = make_case(B=B, gamma=g, V=V, dtype=dtype, sigma=0.5, seed=77, oracle=oracle_mod)
    o0 = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"])
    # put u on / next to the accept boundary for a third of the positions (before and after the first reject)
    ua = case["u_accept"].clone()
    ratio = torch.from_numpy(o0.p_tok / np.maximum(o0.q_tok, 1e-30)).float()
    rng = np.random.RandomState(3)
    for b in range(B):
        for i in range(g):
            k = rng.randint(6)
            if k < 3 and 0.0 < float(ratio[b, i]) < 1.0:
                ua[b, i] = float(ratio[b, i]) * (1.0 + (k - 1) * 1e-4)
    case["u_accept"] = ua
    kw = dict(greedy=greedy)
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], ua, case["u_sample"], **kw)
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
    # (the fused and the split tails of the three-launch pipeline: same row kernel, so the fast p/q values of the
    # sure positions coincide bit for bit; the one-launch cluster path of small batches sums them in another order)
    assert lib.specdec_set_option(b"small_b", 0) == 0
    r1 = sd.fused_verify(*args, **kw)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"tail_slots", 0) == 0
    r1b = sd.fused_verify(*args, **kw)   # atomics + counters exchange (tail_fused_kernel)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"tail_slots", 1) == 0
    for a in ("n_accepted", "next_token", "accept_mask", "packed", "next_prob", "p_tok", "q_tok", "first_stop"):
        assert torch.equal(getattr(r1, a), getattr(r1b, a)), a
    assert lib.specdec_set_option(b"no_fused_tail", 1) == 0
    try:
        r2 = sd.fused_verify(*args, **kw)
        torch.cuda.synchronize()
    finally:
        lib.specdec_set_option(b"no_fused_tail", 0)
    _assert_same(o, r1)
    _assert_same(o, r2)
    for a, b_ in ((r1.n_accepted, r2.n_accepted), (r1.next_token, r2.next_token), (r1.accept_mask, r2.accept_mask),
                  (r1.packed, r2.packed), (r1.next_prob, r2.next_prob)):
        assert torch.equal(a, b_)
    # p_tok / q_tok are exact (== oracle bits) up to and including the deciding position in both pipelines;
    # behind it the split pipeline may have made one more position exact, the other reports the 1e-6 fast value
    n = r1.n_accepted.long().clamp(max=g - 1)
    idx = torch.arange(g, device="cuda")[None, :] <= n[:, None]
    assert torch.equal(r1.p_tok[idx], r2.p_tok[idx]) and torch.equal(r1.q_tok[idx], r2.q_tok[idx])
    np.testing.assert_allclose(r1.p_tok.cpu().numpy(), r2.p_tok.cpu().numpy(), rtol=1e-5)
    assert 0 < int(o.n_accepted.sum()) < B * g


@pytest.mark.parametrize("dtype", ["bf16", "f32", "f16"])
@pytest.mark.parametrize("scale,V", [(3.0, 128256), (1.0, 128256), (0.3, 128256), (0.02, 128256), (1.0, 50257), (0.5, 151936)])
def test_histogram_nucleus_on_flat_rows(oracle_mod, dtype, scale, V):
    """flat rows (random-init models: the nucleus holds thousands of tokens up to ~90 % of the vocabulary) go
    through nucleus_hist_kernel (radix-select over the value axis by private histograms + one exact sweep);
    the kept sets / probabilities equal the band-search path bit for bit and the verify equals the oracle."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    B, g = 3, 2
    case = make_case(B=B, gamma=g, V=V, dtype=dtype, sigma=0.3 * scale, seed=int(scale * 100) + V % 13, scale=scale)
    gen = torch.Generator().manual_seed(11)
    for mode in ("nucleus0.9", "nucleus0.9_t0.7"):
        m = MODES[mode]
        tok, _ = oracle_mod.sample_rows(case["draft"].float().numpy().reshape(B * g, V),
                                        torch.rand(B * g, generator=gen).numpy(), **m)
        case["draft_tokens"] = torch.from_numpy(tok.reshape(B, g))
        args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
        r1 = sd.fused_verify(*args, **m)
        p1, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
        assert lib.specdec_set_option(b"no_hist_nucleus", 1) == 0
        try:
            r2 = sd.fused_verify(*args, **m)
            p2, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
        finally:
            lib.specdec_set_option(b"no_hist_nucleus", 0)
        assert torch.equal(p1, p2), "kept set / probabilities differ from the band-search path"
        assert torch.equal(r1.n_accepted, r2.n_accepted) and torch.equal(r1.next_token, r2.next_token)
        assert torch.equal(r1.p_tok, r2.p_tok) and torch.equal(r1.q_tok, r2.q_tok)
        kept = (p1 > 0).sum(-1)
        assert int(kept.min()) > 1000  # genuinely large nuclei
        o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
        _assert_same(o, r1)


def test_histogram_nucleus_resolves_flat_rows_without_fallback():
    """perf guard: on synthetic flat rows every row is resolved by nucleus_fast / nucleus_hist on the first
    attempt (a row left to the band-search slow path costs ~100x; results would still be exact)."""
    import ctypes
    import specdec_b200 as sd
    lib = sd._lib.lib()
    st = (ctypes.c_ulonglong * 16)()
    assert lib.specdec_debug_stats(ctypes.cast(st, ctypes.c_void_p), 1) == 0
    gen = torch.Generator(device="cuda").manual_seed(3)
    B, g, V = 32, 4, 128256
    for scale in (3.0, 0.5, 0.02):
        t = (scale * torch.randn(B, g + 1, V, device="cuda", generator=gen)).to(torch.bfloat16)
        d = (t[:, :g].float() + 0.2 * scale * torch.randn(B, g, V, device="cuda", generator=gen)).to(torch.bfloat16)
        toks = torch.randint(V, (B, g), device="cuda", generator=gen)
        sd.fused_verify(t, d, toks, torch.rand(B, g, device="cuda", generator=gen),
                        torch.rand(B, device="cuda", generator=gen), top_p=0.9)
    torch.cuda.synchronize()
    assert lib.specdec_debug_stats(ctypes.cast(st, ctypes.c_void_p), 1) == 0
    s = list(st)
    assert s[8] == 3 * B * (2 * g + 1), s      # every row resolved by the histogram kernel
    assert s[7] == 0, s                        # nothing left to the slow path
    assert sum(s[0:7]) <= 0.05 * s[8], s       # retried attempts (wider slack) stay rare


@pytest.mark.parametrize("B", [8, 160])
def test_verify_step_is_cuda_graph_capturable(oracle_mod, B):
    """the whole verify step (memset, row kernel(s), plan, fused tail; for B >= 128 bf16 the two-chunk fork/join
    on the library's auxiliary stream) can be captured into a CUDA graph and replayed on new logits."""
    import specdec_b200 as sd
    V, g = 32000, 3
    c1 = make_case(B=B, gamma=g, V=V, dtype="bf16", sigma=0.5, seed=61, oracle=oracle_mod)
    c2 = make_case(B=B, gamma=g, V=V, dtype="bf16", sigma=0.8, seed=62, oracle=oracle_mod)
    keys = ("target", "draft", "draft_tokens", "u_accept", "u_sample")
    static = [c1[k].cuda().clone() for k in keys]
    sd.fused_verify(*static)  # eager warm-up (creates the library's stream / events, sizes the workspace)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        sd.fused_verify(*static)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        r = sd.fused_verify(*static)
    for case in (c2, c1):
        for dst, k in zip(static, keys):
            dst.copy_(case[k].cuda())
        graph.replay()
        torch.cuda.synchronize()
        o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"])
        _assert_same(o, r)


@pytest.mark.parametrize("kind,scale", [("peaked", 3.0), ("randn", 3.0), ("randn", 0.1)])
def test_tma_prepass_of_top_p_rows_changes_nothing(oracle_mod, kind, scale):
    """optionally, top-p rows get their max / MUFU mass / candidate threshold from a launch of the TMA row pipeline
    (rowfast_tma_kernel<DT, true>) instead of nucleus_fast_kernel's own first sweep: identical kept sets,
    probabilities and decisions, and equal to the oracle."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    B, g, V = 6, 3, 128256
    case = make_case(B=B, gamma=g, V=V, dtype="bf16", sigma=0.3 * scale, seed=91, scale=scale, kind=kind)
    m = MODES["nucleus0.9_t0.7"]
    tok, _ = oracle_mod.sample_rows(case["draft"].float().numpy().reshape(B * g, V),
                                    torch.rand(B * g, generator=torch.Generator().manual_seed(4)).numpy(), **m)
    case["draft_tokens"] = torch.from_numpy(tok.reshape(B, g))
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
    r1 = sd.fused_verify(*args, **m)
    p1, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
    assert lib.specdec_set_option(b"no_tma_nucleus", 0) == 0  # enable the (optional) TMA pre-pass
    try:
        r2 = sd.fused_verify(*args, **m)
        p2, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
    finally:
        lib.specdec_set_option(b"no_tma_nucleus", 1)
    assert torch.equal(p1, p2)
    for a, b_ in ((r1.n_accepted, r2.n_accepted), (r1.next_token, r2.next_token), (r1.p_tok, r2.p_tok), (r1.q_tok, r2.q_tok)):
        assert torch.equal(a, b_)
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
    _assert_same(o, r1)


@pytest.mark.parametrize("mode", ["multinomial", "nucleus0.9", "topk50"])
def test_masked_out_vocabulary_regions(oracle_mod, mode):
    """large contiguous / strided regions of -inf logits (constrained decoding): whole threads, warps and TMA
    stages of the row kernels see nothing but -inf."""
    B, g, V = 4, 3, 128256
    case = make_case(B=B, gamma=g, V=V, dtype="bf16", sigma=0.5, seed=123)
    t, d = case["target"].float(), case["draft"].float()
    t[:, :, 1000:70000] = float("-inf"); d[:, :, 1000:70000] = float("-inf")      # a 69k-token hole
    t[1, :, 8 * 17::8 * 256] = 5.0                                                 # (keeps a few tokens of one thread's stride alive)
    for k in range(8):                                                             # every vector of consumer thread 5 is -inf
        t[2, :, (5 * 8 + k)::256 * 8] = float("-inf"); d[2, :, (5 * 8 + k)::256 * 8] = float("-inf")
    t[3, :, 100000:] = float("-inf"); d[3, :, 90000:] = float("-inf")              # different supports for p and q
    case["target"], case["draft"] = t.to(torch.bfloat16), d.to(torch.bfloat16)
    m = MODES[mode]
    tok, _ = oracle_mod.sample_rows(case["draft"].float().numpy().reshape(B * g, V),
                                    torch.rand(B * g, generator=torch.Generator().manual_seed(9)).numpy(), **m)
    case["draft_tokens"] = torch.from_numpy(tok.reshape(B, g))
    o, r = _run_both(oracle_mod, case, mode)
    _assert_same(o, r)
    assert np.isfinite(r.p_tok.cpu().numpy()).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_ngram_greedy_tma_argmax_edge_rows(oracle_mod, dtype):
    """greedy n-gram verify on 16-bit aligned rows takes max / sum / FIRST index of the maximum from the TMA row
    pipeline (rowfast_tma_kernel<DT, 2>): maxima at the row ends, tied maxima in different threads' stripes, rows
    whose maximum is too small for the 16-bit spacing rule (exact fallback), -inf regions."""
    import specdec_b200 as sd
    B, g, V = 6, 3, 128256
    gen = torch.Generator().manual_seed(21)
    t = 2.0 * torch.randn(B, g + 1, V, generator=gen)
    t[0, :, 0] = 30.0                                  # maximum at index 0
    t[1, :, V - 1] = 30.0                              # maximum at the last index
    t[2, :, 70001] = 25.0; t[2, :, 1234] = 25.0; t[2, :, 99999] = 25.0   # tied maxima: the first index wins
    t[3] = 0.0                                         # every logit tied at 0 (maximum 0: spacing rule => exact fallback)
    t[4] = 1e-4 * torch.randn(g + 1, V, generator=gen)  # tiny logits: spacing below the resolution => exact fallback
    t[5, :, 500:120000] = float("-inf")                # masked-out region
    tt = t.to(dtype)
    toks = tt[:, :g].float().argmax(-1)
    toks[1, 2] = 7; toks[5, 0] = 600                   # a rejected draft; a draft inside the masked region
    ua = torch.zeros(B, g); us = torch.zeros(B)
    o = oracle_mod.verify(tt, None, toks, ua, us, greedy=True, flags=F_NGRAM)
    lib = sd._lib.lib()
    assert lib.specdec_set_option(b"tma_ngram", 0) == 0
    try:
        r2 = sd.fused_verify(tt.cuda(), None, toks.cuda(), ua.cuda(), us.cuda(), greedy=True, flags=F_NGRAM)  # LDG arg-max kernel
        assert lib.specdec_set_option(b"tma_ngram", 1) == 0
        r = sd.fused_verify(tt.cuda(), None, toks.cuda(), ua.cuda(), us.cuda(), greedy=True, flags=F_NGRAM)
    finally:
        lib.specdec_set_option(b"tma_ngram", TMA_NGRAM_DEFAULT)
    _assert_same(o, r, ngram=True)
    _assert_same(o, r2, ngram=True)
    assert int(r.next_token[2]) in (1234,) or int(o.n_accepted[2]) < g  # (bonus row of sequence 2: first tied index)
    assert torch.equal(r.n_accepted, r2.n_accepted) and torch.equal(r.next_token, r2.next_token)


# ---- the persistent single-launch path (csrc/mega.cuh, option "mega"=1; off by default) ----
@pytest.mark.parametrize("B,gamma,V,dtype,greedy,flags", [
    (256, 4, 128256, "bf16", False, 0), (64, 4, 128256, "bf16", True, 0), (7, 3, 32000, "f16", False, F_SKIP),
    (33, 6, 50264, "bf16", False, F_BATCHED | F_NO_BONUS | F_FALLBACK), (1, 8, 128256, "bf16", False, 0),
    (300, 2, 4096, "bf16", False, F_NO_BONUS)])
def test_megakernel_matches_oracle_and_three_launch_pipeline(oracle_mod, B, gamma, V, dtype, greedy, flags):
    """one cooperative launch (row-slice streaming CTAs + exact-item CTAs exchanging self-validating words): identical
    outputs to the three-launch pipeline and to the oracle, incl. accept tests inside the fast path's margin."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    case = make_case(B=B, gamma=gamma, V=V, dtype=dtype, sigma=0.5, seed=91 + B, oracle=oracle_mod)
    o0 = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"])
    ua = case["u_accept"].clone()
    ratio = torch.from_numpy(o0.p_tok / np.maximum(o0.q_tok, 1e-30)).float()
    rng = np.random.RandomState(5)
    for b in range(0, B, 3):  # u on / next to the accept boundary: the in-kernel exact route for ambiguous positions
        i = int(rng.randint(gamma))
        if 0.0 < float(ratio[b, i]) < 1.0:
            ua[b, i] = float(ratio[b, i]) * (1.0 + (int(rng.randint(3)) - 1) * 1e-4)
    case["u_accept"] = ua
    tgt = case["target"][:, :-1] if (flags & F_NO_BONUS) else case["target"]
    kw = dict(greedy=greedy, flags=flags, stop_tokens=[int(case["draft_tokens"][0, 0])])
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], ua, case["u_sample"],
                          greedy=greedy, flags=flags, stop_tokens=kw["stop_tokens"])
    args = [tgt.cuda(), case["draft"].cuda(), case["draft_tokens"].cuda(), ua.cuda(), case["u_sample"].cuda()]
    r0 = sd.fused_verify(*args, **kw)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"mega", 1) == 0
    rs = [sd.fused_verify(*args, **kw) for _ in range(3)]  # (workspace reuse across back-to-back launches)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"mega", 0) == 0
    for r1 in rs:
        _assert_same(o, r1)
        for a, b_ in ((r0.n_accepted, r1.n_accepted), (r0.next_token, r1.next_token), (r0.accept_mask, r1.accept_mask),
                      (r0.packed, r1.packed), (r0.next_prob, r1.next_prob), (r0.first_stop, r1.first_stop)):
            assert torch.equal(a, b_)


@pytest.mark.parametrize("mode", ["topk50", "topk50_p0.9", "nucleus0.9", "greedy_topk"])
@pytest.mark.parametrize("flags,sigma", [(0, 0.5), (F_SKIP, 0.5), (F_BATCHED | F_NO_BONUS | F_FALLBACK, 0.0), (0, 0.0), (F_NO_BONUS, 2.0)])
def test_kept_list_sampling_equals_row_sweep(oracle_mod, mode, flags, sigma):
    """masked modes with small kept sets (top-k, top-k + top-p, LLM-like top-p): the next token is drawn from the
    kept-token lists the selection kernels emit (sample_lists_kernel) -- bit-identical to the sweep over the row
    ("no_klist"=1) and to the oracle, incl. greedy, the zero-residual fallback (p == q) and all flag variants."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    B, g, V = 37, 4, 128256
    m = dict(MODES["topk50"], greedy=True) if mode == "greedy_topk" else MODES[mode]
    case = make_case(B=B, gamma=g, V=V, dtype="bf16", sigma=sigma, seed=66, kind="peaked")
    gq = torch.Generator().manual_seed(6)
    t = case["target"].float() * 0.5
    idx = torch.randint(V, (B, g + 1, 24), generator=gq)
    t.scatter_(2, idx, 12.0 + 8.0 * torch.rand(B, g + 1, 24, generator=gq))
    t[0, :, 100:130] = 19.0   # a 30-way tie at the top: the top-k boundary falls inside a tie group
    t[1, :, 5000:5080] = 21.0  # an 80-way tie: more kept tokens than a list holds -> that sequence takes the sweep
    d = t[:, :g] + sigma * torch.randn(B, g, V, generator=gq)
    case["target"], case["draft"] = t.to(torch.bfloat16), d.to(torch.bfloat16)
    tok, _ = oracle_mod.sample_rows(case["draft"].float().numpy().reshape(B * g, V),
                                    torch.rand(B * g, generator=torch.Generator().manual_seed(4)).numpy(), **m)
    case["draft_tokens"] = torch.from_numpy(tok.reshape(B, g))
    tgt = case["target"][:, :-1] if (flags & F_NO_BONUS) else case["target"]
    args = [tgt.cuda(), case["draft"].cuda(), case["draft_tokens"].cuda(), case["u_accept"].cuda(), case["u_sample"].cuda()]
    r1 = sd.fused_verify(*args, flags=flags, **m)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"no_klist", 1) == 0
    r2 = sd.fused_verify(*args, flags=flags, **m)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"no_klist", 0) == 0
    assert lib.specdec_set_option(b"split_lists", 1) == 0  # plan and list draw as two launches instead of plan_lists_kernel
    r3 = sd.fused_verify(*args, flags=flags, **m)
    torch.cuda.synchronize()
    assert lib.specdec_set_option(b"split_lists", 0) == 0
    for rr in (r2, r3):
        for a in ("n_accepted", "next_token", "p_tok", "q_tok", "accept_mask", "packed", "next_prob"):
            assert torch.equal(getattr(r1, a), getattr(rr, a)), a
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], flags=flags, **m)
    _assert_same(o, r1)
    # the drafter-side sampling op takes the same route
    tk1, pt1 = sd.sample_rows(case["draft"].cuda().reshape(B * g, V), None, seed=5, **m)
    assert lib.specdec_set_option(b"no_klist", 1) == 0
    tk2, pt2 = sd.sample_rows(case["draft"].cuda().reshape(B * g, V), None, seed=5, **m)
    assert lib.specdec_set_option(b"no_klist", 0) == 0
    assert torch.equal(tk1, tk2) and torch.equal(pt1, pt2)


@pytest.mark.parametrize("mode", ["topk50", "topk50_p0.9", "nucleus0.9", "nucleus0.9_t0.7"])
@pytest.mark.parametrize("kind,dtype,V", [("peaked", "bf16", 128256), ("randn", "bf16", 128256), ("peaked", "f16", 50264),
                                          ("randn", "bf16", 1000)])
def test_streamed_selection_kernel_equals_exact_kernels(oracle_mod, mode, kind, dtype, V):
    """masked modes on 16-bit rows: rowsel_tma_kernel (TMA-streamed pass, threshold among the slice maxima, gather of
    the hot slices from L2, exact selection by a selector warp) gives bit-identical kept sets / probabilities /
    decisions to the plain-load kernels ("no_rowsel"=1) and equals the oracle; rows it cannot resolve (flat top-p rows,
    candidate overflow) fall through to those kernels."""
    import specdec_b200 as sd
    lib = sd._lib.lib()
    B, g = 10, 3
    case = make_case(B=B, gamma=g, V=V, dtype=dtype, sigma=0.4, seed=55, kind=kind)
    if kind == "peaked":  # LLM-like: a handful of tokens carry almost all the mass
        gq = torch.Generator().manual_seed(6)
        t = case["target"].float() * 0.5
        idx = torch.randint(V, (B, g + 1, 24), generator=gq)
        t.scatter_(2, idx, 12.0 + 8.0 * torch.rand(B, g + 1, 24, generator=gq))
        t[0, :, 100:140] = 19.0     # a 40-way tie at the top (the top-k boundary inside a tie group)
        t[1, :, 0:3000] = 18.0      # a 3000-way tie: more candidates than the buffer holds -> falls through
        t[2, :, 8 * 7::8 * 256] = 17.0  # every candidate in ONE thread's slice
        d = t[:, :g] + 0.4 * torch.randn(B, g, V, generator=gq)
        case["target"], case["draft"] = t.to(case["target"].dtype), d.to(case["target"].dtype)
    m = MODES[mode]
    tok, _ = oracle_mod.sample_rows(case["draft"].float().numpy().reshape(B * g, V),
                                    torch.rand(B * g, generator=torch.Generator().manual_seed(4)).numpy(), **m)
    case["draft_tokens"] = torch.from_numpy(tok.reshape(B, g))
    args = [case[k].cuda() for k in ("target", "draft", "draft_tokens", "u_accept", "u_sample")]
    r1 = sd.fused_verify(*args, **m)
    p1, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
    assert lib.specdec_set_option(b"no_rowsel", 1) == 0
    r2 = sd.fused_verify(*args, **m)
    p2, _ = sd.process_probs(case["target"].cuda(), m["temperature"], m["top_k"], m["top_p"])
    assert lib.specdec_set_option(b"no_rowsel", 0) == 0
    assert torch.equal(p1, p2), "kept set / probabilities differ from the plain-load kernels"
    for a, b_ in ((r1.n_accepted, r2.n_accepted), (r1.next_token, r2.next_token), (r1.p_tok, r2.p_tok), (r1.q_tok, r2.q_tok),
                  (r1.accept_mask, r2.accept_mask)):
        assert torch.equal(a, b_)
    o = oracle_mod.verify(case["target"], case["draft"], case["draft_tokens"], case["u_accept"], case["u_sample"], **m)
    _assert_same(o, r1)
