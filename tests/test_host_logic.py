"""CPU tests: canonical primitives, C-ABI surface, host-side logic, world_size-2 gloo path."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cexp2_accuracy_and_edges(oracle_mod):
    t = np.linspace(-60, 1, 400001).astype(np.float32)
    e = oracle_mod.cexp2(t)
    rel = np.abs(e.astype(np.float64) / np.exp2(t.astype(np.float64)) - 1)
    assert rel.max() < 2.5e-7
    assert oracle_mod.cexp2(np.array([0.0], np.float32))[0] == 1.0
    assert oracle_mod.cexp2(np.array([-1.0, -2.0, -10.0], np.float32)).tolist() == [0.5, 0.25, 2.0 ** -10]
    tiny = oracle_mod.cexp2(np.array([-np.inf, -1e30, -200.0], np.float32))
    assert (tiny > 0).all() and (tiny < 1e-37).all()  # clamped at 2^-125: contributes 0 at 2^-40 resolution


def test_philox_known_answer(oracle_mod):
    # Random123 known-answer test: Philox4x32-10, counter = key = 0 -> first word 0x6627e8d5
    ua, us = oracle_mod.philox_uniform(0, 0, 0, 1, 1)
    assert ua[0, 0] == np.float32((0x6627E8D5 >> 8) * 2.0 ** -24)
    ua2, _ = oracle_mod.philox_uniform(0, 0, 0, 4, 3)
    ua3, _ = oracle_mod.philox_uniform(0, 0, 2, 2, 3)
    assert np.array_equal(ua2[2:], ua3)  # keyed by the GLOBAL sequence id
    assert ua2.min() >= 0 and ua2.max() < 1


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "specdec_b200.h")).read()
    declared = set(re.findall(r"SPECDEC_API\s+[\w\s\*]+?\b(specdec_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    so = os.path.join(ROOT, "speculative-decoding_b200", "libspecdec_b200.so")
    assert os.path.exists(so), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(so)  # loading needs libcudart only, not a GPU
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/specdec_b200.h but not exported"
    from specdec_b200 import _lib
    assert set(_lib.exported_symbols()) == declared
    lib.specdec_version.restype = ctypes.c_int
    assert lib.specdec_version() >= 100
    lib.specdec_error_string.restype = ctypes.c_char_p
    assert lib.specdec_error_string(-2) == b"workspace too small"
    lib.specdec_verify_workspace_bytes.restype = ctypes.c_size_t
    assert lib.specdec_verify_workspace_bytes(256, 4, 128256) > 256 * 9 * 32


def test_sass_is_sm100a():
    so = os.path.join(ROOT, "speculative-decoding_b200", "libspecdec_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_ops_refuse_cpu_tensors_and_no_fallback():
    import specdec_b200 as sd
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        sd.process_probs(torch.zeros(2, 16))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        sd.fused_verify(torch.zeros(1, 3, 16), torch.zeros(1, 2, 16), torch.zeros(1, 2, dtype=torch.long))
    with pytest.raises(RuntimeError):
        sd.NGramStorage(3, 100, device="cpu")
    src = ""
    pk = os.path.join(ROOT, "speculative-decoding_b200")
    for f in os.listdir(pk):
        if f.endswith(".py"):
            src += open(os.path.join(pk, f)).read()
    assert "oracle" not in src.replace("oracle-reproducible", ""), "product code must not touch oracle/"


def test_processor_constructor_contract():
    import specdec_b200 as sd
    assert sd.GreedyProcessor().temperature == 1 and sd.GreedyProcessor().greedy
    p = sd.TopKNucleusProcessor(0.7, 50, 0.9)
    assert (p.temperature, p.top_k, p.top_p) == (0.7, 50, 0.9)
    assert sd.NucleusProcessor(1.0, 0.9).fused_params() == dict(temperature=1.0, top_k=0, top_p=0.9, greedy=False)
    assert issubclass(sd.TopKProcessor, sd.MultinomialProcessor) and issubclass(sd.MultinomialProcessor, sd.LogitsProcessor)
    with pytest.raises(ValueError, match="Unsupported cache type"):
        sd.prune_cache(object(), 1)
    assert sd.prune_cache(None, 3) is None
    t = torch.arange(2 * 3 * 5 * 4.0).reshape(2, 3, 5, 4)
    pr = sd.prune_tuple_cache(((t, t), None, (t,)), 2)
    assert pr[0][0].shape == (2, 3, 3, 4) and pr[1] is None and pr[0][0].data_ptr() == t.data_ptr()
    with pytest.raises(AssertionError):
        sd.ngram_storage.INgramStorage.__init__(object.__new__(sd.NGramStorage), 1, 10)


def test_shard_and_pack_roundtrip():
    from specdec_b200 import dist as sdd
    tot = 0
    for world in (1, 2, 3, 8):
        spans = [sdd.shard_range(259, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 259
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        tot += 1
    n = torch.tensor([0, 2, 4, 1])
    toks = torch.arange(16).reshape(4, 4) + 100
    x = torch.tensor([7, 8, 9, -1])
    pk = sdd.pack_results(n, toks, x)
    assert pk.tolist() == [[0, 7, -1, -1, -1, -1], [2, 104, 105, 8, -1, -1], [4, 108, 109, 110, 111, 9],
                           [1, 112, -1, -1, -1, -1]]
    nn, tt = sdd.unpack_results(pk)
    assert nn.tolist() == [0, 2, 4, 1] and tt[2].tolist() == [108, 109, 110, 111, 9]


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from specdec_b200 import dist as sdd
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
for total in (8, 7):
    lo, hi = sdd.shard_range(total, rank, world)
    g = torch.Generator().manual_seed(0)
    n_all = torch.randint(0, 5, (total,), generator=g)
    toks_all = torch.randint(0, 1000, (total, 4), generator=g)
    x_all = torch.randint(0, 1000, (total,), generator=g)
    full = sdd.pack_results(n_all, toks_all, x_all)
    mine = sdd.pack_results(n_all[lo:hi], toks_all[lo:hi], x_all[lo:hi])
    got = sdd.all_gather_packed(mine, total)
    assert torch.equal(got, full), (rank, total)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_all_gather_packed_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29671")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0 and "ok" in out, out


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints the contract's JSON line (tiny shape here): the unmodified reference from
    oracle/_ref when that copy exists (build container, GPU box), else the torch port."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--V", "4096", "--cpu-sample-B", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["value"] > 0
    from oracle import ref_arm
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_arm.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["steps"] == 1 and set(d["config"]) == {"workload", "l2", "parallelism"}
    assert abs(d["ms_per_step"] - d["ms_per_sample_step"] * 256 / 2) < 1e-6 * d["ms_per_step"]


def test_oracle_verify_semantics_small(oracle_mod):
    """hand-checkable cases of the verify semantics (SURVEY 8a spec items 3-5)."""
    V = 4
    t = np.log(np.array([[[0.1, 0.2, 0.3, 0.4], [0.25, 0.25, 0.25, 0.25], [0.7, 0.1, 0.1, 0.1]]], np.float32))
    d = np.log(np.array([[[0.4, 0.3, 0.2, 0.1], [0.25, 0.25, 0.25, 0.25]]], np.float32))
    # token 0: p/q = 0.25 ; token 3: p/q = 4 (always accepted)
    o = oracle_mod.verify(t, d, [[0, 1]], [[0.2, 0.999]], [0.5])
    assert o.n_accepted[0] == 2 and o.accept_mask.tolist() == [[1, 1]]
    assert o.next_token[0] == 0  # bonus row [0.7,...] at u = 0.5
    np.testing.assert_allclose(o.p_tok[0], [0.1, 0.25], rtol=1e-6)
    np.testing.assert_allclose(o.q_tok[0], [0.4, 0.25], rtol=1e-6)
    o = oracle_mod.verify(t, d, [[0, 1]], [[0.3, 0.0]], [0.5])
    assert o.n_accepted[0] == 0 and o.accept_mask.tolist() == [[0, 1]]
    # residual max(0,p-q) = [0,0,.1,.3]/.4 -> cdf .25 | 1.0 ; u=.5 -> token 3, u=.2 -> token 2
    assert o.next_token[0] == 3
    assert oracle_mod.verify(t, d, [[0, 1]], [[0.3, 0.0]], [0.2]).next_token[0] == 2
    assert oracle_mod.verify(t, d, [[0, 1]], [[0.3, 0.0]], [0.2], greedy=True).next_token[0] == 3
    # skip_sample_adjustment: sample p[0] = [.1,.2,.3,.4] at u = .2 -> token 1
    assert oracle_mod.verify(t, d, [[0, 1]], [[0.3, 0.0]], [0.2], flags=4).next_token[0] == 1
    # batched rule: strict '<', no bonus
    o = oracle_mod.verify(t, d, [[3, 1]], [[0.999, 0.999]], [0.5], flags=1 | 2 | 16)
    assert o.n_accepted[0] == 2 and o.next_token[0] == -1
    # stop tokens: first accepted draft that is a stop token
    o = oracle_mod.verify(t, d, [[3, 1]], [[0.1, 0.1]], [0.5], stop_tokens=[1, 2])
    assert o.first_stop[0] == 1
    # top-k=1 on the drafter row makes q_tok = 0 for other tokens -> p/q = inf -> accept
    o = oracle_mod.verify(t, d, [[3, 1]], [[0.99, 0.99]], [0.5], top_k=1)
    assert o.q_tok[0, 0] == 0 and o.accept_mask[0, 0] == 1


def test_every_library_option_is_documented_in_the_header():
    """specdec_set_option names accepted by csrc/verify.cu == names documented in include/specdec_b200.h."""
    src = open(os.path.join(ROOT, "speculative-decoding_b200", "csrc", "verify.cu")).read()
    body = src[src.index("int specdec_set_option("):]
    body = body[:body.index("\n}\n")]
    accepted = set(re.findall(r'strcmp\(name, "(\w+)"\)', body))
    hdr = open(os.path.join(ROOT, "include", "specdec_b200.h")).read()
    doc = hdr[:hdr.index("SPECDEC_API int specdec_set_option")]
    documented = set(re.findall(r'"(\w+)"(?:=\w+)?', doc[doc.rindex("/*"):]))
    internal = {"mega_r", "mega_unit", "mega_spc", "mega_keep_l2", "mega_dbg", "rowsel_probe", "chunk0_pct"}  # tuning probes
    assert accepted - internal <= documented, sorted(accepted - internal - documented)
    lib = ctypes.CDLL(os.path.join(ROOT, "speculative-decoding_b200", "libspecdec_b200.so"))
    lib.specdec_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    assert lib.specdec_set_option(b"no_such_option", 1) != 0 and lib.specdec_set_option(b"reset", 1) == 0
