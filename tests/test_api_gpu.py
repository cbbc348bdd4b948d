"""GPU tests of the remaining C-ABI entry points and of the drop-in Python API."""
import numpy as np
import pytest
import torch

from cases import MODES, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_process_probs_matches_oracle(oracle_mod, mode, dtype):
    import specdec_b200 as sd
    m = MODES[mode]
    z = make_case(B=3, gamma=2, V=32000, dtype=dtype, seed=4)["target"]
    ref, st = oracle_mod.process_probs(z, temperature=m["temperature"], top_k=m["top_k"], top_p=m["top_p"])
    probs, stats = sd.process_probs(z.cuda(), m["temperature"], m["top_k"], m["top_p"])
    assert np.array_equal(probs.cpu().numpy(), ref)  # same canonical arithmetic => identical bits
    assert np.array_equal(stats[:, 0].cpu().numpy(), st["m"]) and np.array_equal(stats[:, 1].cpu().numpy(), st["S32"])
    np.testing.assert_allclose(probs.sum(-1).cpu().numpy(), 1.0, rtol=1e-5)
    # against torch's own softmax (the reference's arithmetic): 1e-5 relative on everything that matters
    if m["top_k"] == 0 and m["top_p"] >= 1.0:
        t = torch.softmax(z.cuda().float() / m["temperature"], -1)
        big = t > 1e-12
        rel = ((probs - t).abs() / t)[big].max().item()
        assert rel < 2e-5


def test_processor_classes_api(oracle_mod):
    import specdec_b200 as sd
    z = make_case(B=2, gamma=1, V=5000, dtype="bf16", seed=8)["target"].cuda()
    for proc, kw in [(sd.GreedyProcessor(), dict(temperature=1.0)), (sd.MultinomialProcessor(0.7), dict(temperature=0.7)),
                     (sd.TopKProcessor(0.7, 50), dict(temperature=0.7, top_k=50)),
                     (sd.NucleusProcessor(1.0, 0.9), dict(temperature=1.0, top_p=0.9)),
                     (sd.TopKNucleusProcessor(0.7, 50, 0.9), dict(temperature=0.7, top_k=50, top_p=0.9))]:
        zc = z.clone()
        p = proc(zc)
        assert torch.equal(zc, z), "logits must stay read-only"
        assert p.shape == z.shape and p.dtype == z.dtype
        ref, _ = oracle_mod.process_probs(z, **kw)
        assert torch.equal(p.cpu(), torch.from_numpy(ref).to(z.dtype))
        s = proc.sample(p)
        assert s.shape == (*z.shape[:-1], 1) and s.dtype == torch.int64
        masked = proc._process(z)
        assert bool(((masked <= -1e19) == (torch.from_numpy(ref).cuda() == 0)).all()) or not kw.get("top_k") and not kw.get("top_p")
    g = sd.GreedyProcessor()
    assert torch.equal(g.sample(g(z)).squeeze(-1).cpu(), torch.from_numpy(oracle_mod.sample_probs(g(z).float().cpu().numpy(), None, greedy=True)).reshape(2, 2))


@pytest.mark.parametrize("mode", ["greedy", "multinomial", "topk50_p0.9"])
def test_sample_rows_and_probs(oracle_mod, mode):
    import specdec_b200 as sd
    m = MODES[mode]
    z = make_case(B=8, gamma=2, V=32000, dtype="bf16", seed=6)["target"].reshape(-1, 32000)
    u = torch.rand(z.shape[0], generator=torch.Generator().manual_seed(1))
    tok, ptok = sd.sample_rows(z.cuda(), u.cuda(), **m)
    otok, optok = oracle_mod.sample_rows(z, u.numpy(), **m)
    assert np.array_equal(tok.cpu().numpy(), otok) and np.array_equal(ptok.cpu().numpy(), optok)
    probs, _ = oracle_mod.process_probs(z, temperature=m["temperature"], top_k=m["top_k"], top_p=m["top_p"])
    t2 = sd.sample_probs(torch.from_numpy(probs).cuda(), u.cuda(), greedy=m["greedy"])
    assert np.array_equal(t2.cpu().numpy(), oracle_mod.sample_probs(probs, u.numpy(), greedy=m["greedy"]))


def test_residual_resample_chi_square(oracle_mod):
    """Distribution test (north_star): tokens emitted after a rejection follow norm(max(0,p-q))."""
    import specdec_b200 as sd
    from scipy import stats
    V, N = 64, 20000
    gen = torch.Generator().manual_seed(3)
    t = 1.5 * torch.randn(1, 2, V, generator=gen)
    d = 1.5 * torch.randn(1, 1, V, generator=gen)
    p = torch.softmax(t[0, 0], -1); q = torch.softmax(d[0, 0], -1)
    tok = int(torch.argmax(q / p))  # a draft that is (almost) always rejected
    r = sd.fused_verify(t.cuda().expand(N, 2, V), d.cuda().expand(N, 1, V), torch.full((N, 1), tok).cuda(), None, None,
                        seed=99, offset=0)
    rej = r.n_accepted.cpu() == 0
    assert rej.float().mean() > 0.5
    xs = r.next_token.cpu()[rej].numpy()
    resid = torch.clamp(p - q, min=0); resid = (resid / resid.sum()).numpy()
    obs = np.bincount(xs, minlength=V).astype(np.float64)
    keep = resid * len(xs) >= 5
    chi, pval = stats.chisquare(np.append(obs[keep], obs[~keep].sum()),
                                np.append(resid[keep], resid[~keep].sum()) * len(xs))
    assert pval > 1e-3, (chi, pval)
    # acceptance frequency ~ min(1, p/q)
    acc_rate = 1.0 - rej.float().mean().item()
    assert abs(acc_rate - min(1.0, float(p[tok] / q[tok]))) < 0.02


def test_prune_kv_matches_reference_view_semantics(oracle_mod):
    import specdec_b200 as sd
    B, H, S, D = 5, 3, 40, 16
    gen = torch.Generator().manual_seed(0)
    tensors = [torch.randn(B, H, S, D, generator=gen).to(torch.bfloat16) for _ in range(4)]
    lens = torch.tensor([40, 17, 5, 1, 0], dtype=torch.int32)
    lens0 = lens.clone()  # (the oracle updates the array it is given in place)
    disc = torch.tensor([3, 5, 5, 4, 2], dtype=torch.int32)
    dev = [t.cuda() for t in tensors]
    dl = lens.cuda()
    sd.prune_kv(dev, dl, disc.cuda(), True)
    exp = [t.clone().view(torch.int16).numpy() for t in tensors]
    new = oracle_mod.prune_kv(exp, lens.numpy(), disc.numpy(), True)
    assert np.array_equal(dl.cpu().numpy(), new)
    for a, b in zip(dev, exp):
        assert np.array_equal(a.cpu().view(torch.int16).numpy(), b)
    # the valid prefix equals the reference's per-sequence view tensor[:, :, :-n, :]
    for b in range(B):
        n = int(min(disc[b], lens0[b])); L0 = int(lens0[b])
        if L0 - n > 0:
            ref_view = tensors[0][b:b + 1, :, :L0, :][:, :, :L0 - n, :] if n > 0 else tensors[0][b:b + 1, :, :L0, :]
            assert torch.equal(dev[0][b:b + 1, :, :L0 - n, :].cpu(), ref_view)
    # default rollback = the length vector only (one tiny launch): the tensors are not touched, the valid prefix
    # [0, len_b) is the reference's pruned view; StaticKVCache caches its device pointer table for zero fills
    dev2 = [t.cuda() for t in tensors]
    kv = sd.StaticKVCache(dev2, lens0.cuda())
    kv.rollback(disc.cuda())
    assert np.array_equal(kv.seq_lens.cpu().numpy(), new)
    for a, b in zip(dev2, tensors):
        assert torch.equal(a.cpu(), b)
    kv2 = sd.StaticKVCache([t.cuda() for t in tensors], lens0.cuda())
    kv2.rollback(disc.cuda(), zero_fill=True)
    kv2.rollback(0, zero_fill=True)
    for a, b in zip(kv2.tensors, exp):
        assert np.array_equal(a.cpu().view(torch.int16).numpy(), b)
    tv = kv2.as_tuple_views(1)
    assert tv[0][0].shape == (1, H, 12, D)
    # drop-in prune_cache on tuple caches is the same zero-copy view as the reference
    cache = tuple((t[:1], t[:1]) for t in dev[:2])
    pr = sd.prune_cache(cache, 4)
    assert pr[0][0].shape[2] == S - 4 and pr[0][0].data_ptr() == cache[0][0].data_ptr()
    with pytest.raises(ValueError):
        sd.prune_cache([1, 2], 1)


@pytest.mark.parametrize("one_level", [False, True])
def test_ngram_tables_match_oracle(one_level):
    import specdec_b200 as sd
    from oracle.ngram_oracle import NGramOracle
    rng = np.random.RandomState(0)
    V, n, B = 50, 4, 6
    cls = sd.OneLevelNGramStorage if one_level else sd.NGramStorage
    # (a) shared table, as in the reference; (b) one table per sequence
    for per_seq in (False, True):
        st = cls(n, V, n_tables=B if per_seq else 1, grams_per_table=4096, counts_per_table=8192)
        orc = NGramOracle(n, V, one_level)
        tabs = torch.arange(B, dtype=torch.int32) if per_seq else None
        seqs = rng.randint(0, 6, size=(B, 30))  # tiny alphabet => many repeated contexts
        lens = rng.randint(0, 31, size=B).astype(np.int32); lens[0] = 30; lens[1] = 2
        ids = torch.from_numpy(seqs)
        st.initialize(ids, torch.from_numpy(lens), tabs)
        orc.initialize([seqs[i, :lens[i]] for i in range(B)], None if tabs is None else tabs.tolist())
        for step in range(12):
            nt = rng.randint(0, 6, size=(B, 1 + step % 3))
            st.update(ids, torch.from_numpy(nt), torch.from_numpy(lens), tabs)
            orc.update([seqs[i, :lens[i]] for i in range(B)], nt, None if tabs is None else tabs.tolist())
            fb = rng.randint(0, V, size=(B, 5))
            d, k = st.lookup_chain(ids, 5, torch.from_numpy(lens), tabs, torch.from_numpy(fb))
            od, ok = orc.lookup_chain([seqs[i, :lens[i]] for i in range(B)], 5, None if tabs is None else tabs.tolist(), fb)
            assert d.cpu().tolist() == od and k.cpu().tolist() == ok
            lens = np.maximum(lens - (step % 2), 0).astype(np.int32)
        # has_gram: exact (every counted token, not only the arg-max one), ngram_storage.py:98-106 / :181-193
        for trial in range(60):
            L_ = int(rng.randint(1, 8))
            q = rng.randint(0, 6, size=L_)
            tb = int(rng.randint(0, B)) if per_seq else 0
            assert st.has_gram(torch.from_numpy(q), tb) == orc.has_gram(q, tb), (q, tb)
        assert not st.status()["overflow"]
        st.reset()
        d, k = st.lookup_chain(ids, 2, torch.from_numpy(lens), tabs, torch.zeros(B, 2, dtype=torch.long))
        assert not bool(k.any())


def test_ngram_device_fallback_tokens_and_table_id_validation():
    import specdec_b200 as sd
    V, B = 1000, 16
    ids = torch.randint(0, 50, (B, 12), generator=torch.Generator().manual_seed(0))
    outs = []
    for seed in (5, 5, 6):
        st = sd.NGramStorage(4, V, n_tables=B, grams_per_table=256, counts_per_table=512, seed=seed)
        tabs = torch.arange(B, dtype=torch.int32)
        d1, k1 = st.lookup_chain(ids, 4, table_ids=tabs)   # empty tables: everything is a fallback token
        d2, k2 = st.lookup_chain(ids, 4, table_ids=tabs)
        assert not bool(k1.any()) and int(d1.min()) >= 0 and int(d1.max()) < V
        assert not torch.equal(d1, d2), "every call draws fresh fallback tokens"
        outs.append(d1.cpu())
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])
    st = sd.NGramStorage(4, V, n_tables=4)
    with pytest.raises(ValueError):
        st.lookup_chain(ids, 2, table_ids=torch.arange(B, dtype=torch.int32))  # ids 4..15 are out of range
    with pytest.raises(ValueError):
        sd.NGramStorage(11, V)


def test_missing_library_fails_loudly(monkeypatch):
    import specdec_b200 as sd
    monkeypatch.setattr(sd._lib, "_lib", None)
    monkeypatch.setattr(sd._lib, "LIB_PATH", "/nonexistent/libspecdec_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        sd.process_probs(torch.zeros(1, 8, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        sd.ops.verify_op(torch.zeros(1, 2, 8), torch.zeros(1, 1, 8), torch.zeros(1, 1, dtype=torch.long), None, None,
                         0, 0, 0, 1.0, 0, 1.0, 1, 0, None)


# ---- decode-loop helpers (csrc/engine.cu), device-resident Philox offset, default generator ----
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("V,k", [(32000, 3), (128256, 3), (1000, 8), (5003, 11), (7, 7)])
def test_topk_ids_matches_stable_sort(dtype, V, k):
    """ngram_assisted/ngram_assisted.py:149-155: torch.topk(p, filler_top_k).indices; our order is value desc / index asc."""
    import specdec_b200 as sd
    g = torch.Generator().manual_seed(V + k)
    z = (3 * torch.randn(5, V, generator=g)).to(dtype)
    z[1, : V // 2] = z[1, 0]          # a large tie group
    z[2, V - 1] = 100.0               # maximum at the ragged end
    z[3] = float("-inf"); z[3, V // 3] = 0.5
    ids = sd.ops.topk_ids(z.cuda(), k).cpu()
    want = torch.sort(z.float(), dim=-1, descending=True, stable=True).indices[:, :k]
    assert torch.equal(ids, want)
    # strided view of a [B, L, V] model output
    big = torch.zeros(5, 3, V, dtype=dtype); big[:, 1] = z
    assert torch.equal(sd.ops.topk_ids(big.cuda()[:, 1], k).cpu(), want)


def _writeback_torch(generated, step, g, n, fs, x, finished, n_acc, end_tokens):
    """the vectorised torch statement of engine/infer_engine.py:300-336 this kernel replaced"""
    device = generated.device
    active = ~finished
    n, fs = n.long(), fs.long()
    hit_end = fs >= 0
    acc_cnt = torch.where(hit_end, fs + 1, n)
    rejected = (~hit_end) & (n < g)
    n_acc = n_acc + torch.where(active, acc_cnt, torch.zeros_like(acc_cnt))
    ar = torch.arange(g, device=device)
    cur = generated[:, step:step + g]
    corr = rejected.unsqueeze(1) & (ar.unsqueeze(0) == n.unsqueeze(1))
    cur = torch.where(corr, x.unsqueeze(1), cur)
    tail = (acc_cnt < g).unsqueeze(1) & (ar.unsqueeze(0) >= (acc_cnt + 1).unsqueeze(1))
    cur = torch.where(tail, torch.zeros_like(cur), cur)
    generated = generated.clone()
    generated[:, step:step + g] = torch.where(active.unsqueeze(1), cur, generated[:, step:step + g])
    x_is_end = torch.isin(x, end_tokens) if end_tokens.numel() else torch.zeros_like(rejected)
    finished = finished | (active & (hit_end | (rejected & x_is_end)))
    return generated, finished, n_acc


@pytest.mark.parametrize("step_on_device", [False, True])
def test_batch_writeback_matches_torch_statement(step_on_device):
    import specdec_b200 as sd
    gen = torch.Generator().manual_seed(5)
    B, G, g, step = 300, 24, 6, 12
    generated = torch.randint(1, 1000, (B, G), generator=gen).cuda()
    n = torch.randint(0, g + 1, (B,), generator=gen).int().cuda()
    fs = torch.where(torch.rand(B, generator=gen) < 0.2, torch.randint(0, g, (B,), generator=gen), torch.full((B,), -1)).int().cuda()
    fs = torch.where(fs >= n, torch.full_like(fs, -1), fs)  # a stop can only sit among the accepted drafts
    x = torch.randint(0, 20, (B,), generator=gen).cuda()
    finished = (torch.rand(B, generator=gen) < 0.3).cuda()
    n_acc = torch.randint(0, 50, (B,), generator=gen).cuda()
    end_tokens = torch.tensor([3, 7], dtype=torch.long).cuda()

    class R:
        pass
    r = R(); r.n_accepted, r.first_stop, r.next_token = n, fs, x
    wg, wf, wa = _writeback_torch(generated, step, g, n, fs, x, finished, n_acc, end_tokens)
    g2, f2, a2 = generated.clone(), finished.clone(), n_acc.clone()
    n_active = torch.full((1,), -5, dtype=torch.int32).cuda()
    st = torch.tensor([step], dtype=torch.int64).cuda() if step_on_device else step
    sd.ops.batch_writeback(r, g2, st, g, f2, a2, end_tokens, n_active)
    assert torch.equal(g2, wg) and torch.equal(f2, wf) and torch.equal(a2, wa)
    assert int(n_active[0]) == int((~wf).sum())


def test_device_resident_offset_equals_host_offset_and_advances_in_a_graph(oracle_mod):
    import specdec_b200 as sd
    case = make_case(B=6, gamma=4, V=32000, dtype="bf16", sigma=0.5, seed=3, oracle=oracle_mod, mode="multinomial")
    t, d, tk = case["target"].cuda(), case["draft"].cuda(), case["draft_tokens"].cuda()
    off = torch.zeros(1, dtype=torch.int64).cuda()
    host = [sd.fused_verify(t, d, tk, None, None, seed=99, offset=i) for i in range(4)]
    for i in range(4):
        off.fill_(i)
        r = sd.fused_verify(t, d, tk, None, None, seed=99, offset=off)
        assert torch.equal(r.packed, host[i].packed) and torch.equal(r.accept_mask, host[i].accept_mask)
        z = d.reshape(-1, 32000)
        a, _ = sd.sample_rows(z, None, seed=99, offset=off, lane_id=2)
        b, _ = sd.sample_rows(z, None, seed=99, offset=i, lane_id=2)
        assert torch.equal(a, b)
    # one captured step, replayed: the device word advances inside the graph, the uniforms with it
    off.fill_(0)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        sd.fused_verify(t, d, tk, None, None, seed=99, offset=off)
    torch.cuda.current_stream().wait_stream(side)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        rg = sd.fused_verify(t, d, tk, None, None, seed=99, offset=off)
        off.add_(1)
    for i in range(4):
        gph.replay()
        torch.cuda.synchronize()
        assert torch.equal(rg.packed, host[i].packed), i
    assert int(off[0]) == 4
    assert any(not torch.equal(host[0].packed, host[i].packed) for i in range(1, 4))


def test_default_uniforms_advance_across_calls_and_rekey_on_manual_seed():
    import specdec_b200 as sd
    from specdec_b200 import uniforms as _u
    _u._DEFAULT = None
    torch.manual_seed(1234)
    un = sd.default_uniforms()
    assert un.offset == 0 and un is sd.default_uniforms()
    p = torch.softmax(torch.randn(64, 500, generator=torch.Generator().manual_seed(0)), -1).cuda()
    proc = sd.MultinomialProcessor(1.0)
    a = proc.sample(p)
    b = proc.sample(p)
    assert un.offset == 2 and not torch.equal(a, b), "successive calls must not replay the same uniforms"
    proc2 = sd.NucleusProcessor(1.0, 0.9)   # another processor instance shares the generator
    proc2.sample(p)
    assert un.offset == 3
    torch.manual_seed(77)                   # re-keyed by the user: new stream, offset 0
    un2 = sd.default_uniforms()
    assert un2 is not un and un2.offset == 0 and un2.seed == 77


def test_verify_result_host_is_one_copy(oracle_mod):
    import specdec_b200 as sd
    case = make_case(B=5, gamma=3, V=5000, dtype="f32", sigma=0.5, seed=11, oracle=oracle_mod, mode="multinomial")
    r = sd.fused_verify(case["target"].cuda(), case["draft"].cuda(), case["draft_tokens"].cuda(), None, None, seed=5,
                        stop_tokens=[int(case["draft_tokens"][0, 0])])
    n, x, fs = r.host()
    assert n == r.n_accepted.tolist() and x == r.next_token.tolist() and fs == r.first_stop.tolist()


def test_graphed_verify_replays_match_eager_steps(oracle_mod):
    import specdec_b200 as sd
    case = make_case(B=3, gamma=4, V=32000, dtype="bf16", sigma=0.5, seed=13, oracle=oracle_mod, mode="topk50")
    m = MODES["topk50"]
    t, d, tk = case["target"].cuda(), case["draft"].cuda(), case["draft_tokens"].cuda()
    gv = sd.GraphedVerify(t, d, tk, seed=21, offset0=5, **m)
    for i in range(3):
        r = gv()
        e = sd.fused_verify(t, d, tk, None, None, seed=21, offset=5 + i, **m)
        assert torch.equal(r.packed, e.packed) and torch.equal(r.first_stop, e.first_stop)
    # the graph reads the buffers' current contents
    t2 = torch.roll(t, 1, 0)
    t.copy_(t2)
    r = gv()
    e = sd.fused_verify(t, d, tk, None, None, seed=21, offset=8, **m)
    assert torch.equal(r.packed, e.packed)


@pytest.mark.parametrize("one_level", [False, True])
@pytest.mark.parametrize("with_fill", [False, True])
def test_ngram_update_chain_equals_per_position_updates(one_level, with_fill):
    """ngram_assisted/ngram_assisted.py:149-155: the 2 (n + 1) updates of a step as ONE launch == the reference's sequence
    of update() calls (same order => same arg-max tokens), checked through the oracle dicts and the device tables."""
    import specdec_b200 as sd
    from oracle.ngram_oracle import NGramOracle
    rng = np.random.RandomState(7)
    V, n = 40, 4
    cls = sd.OneLevelNGramStorage if one_level else sd.NGramStorage
    a, b_ = cls(n, V, grams_per_table=4096, counts_per_table=8192), cls(n, V, grams_per_table=4096, counts_per_table=8192)
    orc = NGramOracle(n, V, one_level)
    seq = torch.from_numpy(rng.randint(0, 5, size=64)).cuda()
    for st in (a, b_):
        st.initialize(seq[:20].reshape(1, -1))
    orc.initialize([seq[:20].cpu().numpy()])
    pos = 20
    for step in range(8):
        n1 = int(rng.randint(1, 6))
        fill = torch.from_numpy(rng.randint(0, 5, size=(n1, 3))).cuda() if with_fill else None
        a.update_chain(seq, pos, seq[pos:pos + n1], fill)
        for i in range(n1):
            b_.update(seq[:pos + i].reshape(1, -1), seq[pos + i].reshape(1, 1))
            orc.update([seq[:pos + i].cpu().numpy()], [[int(seq[pos + i])]])
            if with_fill:
                b_.update(seq[:pos + i].reshape(1, -1), fill[i].reshape(1, -1))
                orc.update([seq[:pos + i].cpu().numpy()], [fill[i].cpu().tolist()])
        pos += n1
        fb = torch.zeros(1, 4, dtype=torch.long).cuda()
        for L_ in (pos, pos - 1, 10):
            da, ka = a.lookup_chain(seq[:L_].reshape(1, -1), 4, fallback=fb)
            db, kb = b_.lookup_chain(seq[:L_].reshape(1, -1), 4, fallback=fb)
            od, ok = orc.lookup_chain([seq[:L_].cpu().numpy()], 4, None, fb.cpu().numpy())
            assert torch.equal(da, db) and torch.equal(ka, kb)
            assert da.cpu().tolist() == od and ka.cpu().tolist() == ok
