"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under speculative-decoding_b200/ does.
The arithmetic lives in oracle/specdec_oracle.c (see its header for the reference
file:line each function restates).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libspecdec_oracle.so")

F_ACCEPT_BATCHED = 1
F_NO_BONUS = 2
F_SKIP_ADJUST = 4
F_NGRAM = 8
F_RESID_FALLBACK = 16


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "specdec_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_cexp2.restype = C.c_float
        _lib.oracle_cexp2.argtypes = [C.c_float]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_void_p):
    return None if a is None else a.ctypes.data_as(t)


def _f32(x):
    """Accept numpy / torch (any float dtype) -> C-contiguous float32 numpy (exact for bf16/fp16)."""
    if hasattr(x, "detach"):
        x = x.detach().float().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


@dataclass
class VerifyOut:
    n_accepted: np.ndarray
    next_token: np.ndarray
    accept_mask: np.ndarray
    p_tok: np.ndarray
    q_tok: np.ndarray
    first_stop: np.ndarray


def verify(target_logits, draft_logits, draft_tokens, u_accept, u_sample, *, temperature=1.0, top_k=0,
           top_p=1.0, greedy=False, flags=0, stop_tokens=()):
    """One verify step for B sequences.  target [B,g+1,V], draft [B,g,V] or None (n-gram mode)."""
    t = _f32(target_logits)
    B, g1, V = t.shape
    gamma = g1 - 1
    d = None if draft_logits is None else _f32(draft_logits)
    if d is None:
        assert flags & F_NGRAM
    else:
        assert d.shape == (B, gamma, V)
    tok = np.ascontiguousarray(np.asarray(draft_tokens, dtype=np.int64).reshape(B, gamma))
    ua = np.ascontiguousarray(np.asarray(u_accept, dtype=np.float32).reshape(B, gamma))
    us = np.ascontiguousarray(np.asarray(u_sample, dtype=np.float32).reshape(B))
    stop = np.ascontiguousarray(np.asarray(list(stop_tokens), dtype=np.int64))
    out = VerifyOut(np.zeros(B, np.int32), np.zeros(B, np.int64), np.zeros((B, gamma), np.uint8),
                    np.zeros((B, gamma), np.float32), np.zeros((B, gamma), np.float32), np.zeros(B, np.int32))
    rc = lib().oracle_verify(
        _p(t), _p(d), C.c_int(B), C.c_int(gamma), C.c_int64(V),
        C.c_int64(g1 * V), C.c_int64(V), C.c_int64(gamma * V), C.c_int64(V),
        _p(tok), _p(ua), _p(us), C.c_float(temperature), C.c_int(top_k), C.c_float(top_p),
        C.c_int(1 if greedy else 0), C.c_int(flags), _p(stop), C.c_int(stop.size),
        _p(out.n_accepted), _p(out.next_token), _p(out.accept_mask), _p(out.p_tok), _p(out.q_tok),
        _p(out.first_stop))
    if rc != 0:
        raise ValueError("oracle_verify: draft token out of range")
    return out


def process_probs(logits, *, temperature=1.0, top_k=0, top_p=1.0, want_probs=True):
    """LogitsProcessor.__call__ on [..., V] logits -> (probs or None, stats dict)."""
    z = _f32(logits)
    V = z.shape[-1]
    z2 = z.reshape(-1, V)
    rows = z2.shape[0]
    probs = np.zeros((rows, V), np.float32) if want_probs else None
    st = dict(m=np.zeros(rows, np.float32), S32=np.zeros(rows, np.float32), Sfix=np.zeros(rows, np.uint64),
              cut=np.zeros(rows, np.float32), jcut=np.zeros(rows, np.int64), n_kept=np.zeros(rows, np.int64))
    lib().oracle_process_probs(_p(z2), C.c_int64(rows), C.c_int64(V), C.c_int64(V), C.c_float(temperature),
                               C.c_int(top_k), C.c_float(top_p), _p(probs), _p(st["m"]), _p(st["S32"]),
                               _p(st["Sfix"]), _p(st["cut"]), _p(st["jcut"]), _p(st["n_kept"]))
    if want_probs:
        probs = probs.reshape(z.shape)
    return probs, st


def sample_rows(logits, u, *, temperature=1.0, top_k=0, top_p=1.0, greedy=False):
    z = _f32(logits)
    V = z.shape[-1]
    z2 = z.reshape(-1, V)
    rows = z2.shape[0]
    uu = np.ascontiguousarray(np.asarray(u, dtype=np.float32).reshape(rows))
    tok = np.zeros(rows, np.int64)
    ptok = np.zeros(rows, np.float32)
    lib().oracle_sample_rows(_p(z2), C.c_int64(rows), C.c_int64(V), C.c_int64(V), C.c_float(temperature),
                             C.c_int(top_k), C.c_float(top_p), C.c_int(1 if greedy else 0), _p(uu), _p(tok),
                             _p(ptok))
    return tok, ptok


def sample_probs(probs, u, *, greedy=False):
    p = _f32(probs)
    V = p.shape[-1]
    p2 = p.reshape(-1, V)
    rows = p2.shape[0]
    uu = np.ascontiguousarray(np.asarray(u if u is not None else np.zeros(rows), dtype=np.float32).reshape(rows))
    tok = np.zeros(rows, np.int64)
    lib().oracle_sample_probs(_p(p2), C.c_int64(rows), C.c_int64(V), C.c_int(1 if greedy else 0), _p(uu), _p(tok))
    return tok


def philox_uniform(seed, offset, seq0, B, gamma):
    ua = np.zeros((B, gamma), np.float32)
    us = np.zeros(B, np.float32)
    lib().oracle_philox_uniform(C.c_uint64(seed), C.c_uint64(offset), C.c_int64(seq0), C.c_int(B), C.c_int(gamma),
                                _p(ua), _p(us))
    return ua, us


def cexp2(t):
    t = np.ascontiguousarray(t, dtype=np.float32)
    out = np.empty_like(t)
    lib().oracle_cexp2_array(_p(t), _p(out), C.c_int64(t.size))
    return out


def prune_kv(tensors, seq_lens, discard, zero_fill=True):
    """tensors: list of C-contiguous numpy arrays [B,H,S_max,D] (modified in place); seq_lens int32 in/out."""
    B, H, S, D = tensors[0].shape
    eb = tensors[0].dtype.itemsize
    arr = (C.c_void_p * len(tensors))(*[t.ctypes.data for t in tensors])
    lens = np.ascontiguousarray(seq_lens, dtype=np.int32)
    disc = np.ascontiguousarray(discard, dtype=np.int32)
    lib().oracle_prune_kv(arr, C.c_int(len(tensors)), C.c_int(B), C.c_int(H), C.c_int64(S), C.c_int64(D),
                          C.c_int(eb), _p(lens), _p(disc), C.c_int(1 if zero_fill else 0))
    return lens


def num_threads():
    return lib().oracle_num_threads()


def set_threads(n):
    lib().oracle_set_threads(C.c_int(n))
