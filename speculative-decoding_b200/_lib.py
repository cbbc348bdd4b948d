"""ctypes binding of libspecdec_b200.so (the C ABI declared in include/specdec_b200.h).

There is NO fallback: if the CUDA library is missing, importing the ops raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPECDEC_B200_LIB") or os.path.join(_HERE, "libspecdec_b200.so")  # (override: tuning builds)

F32, BF16, F16 = 0, 1, 2
SAMPLE_GREEDY, SAMPLE_INVCDF = 0, 1
ACCEPT_BATCHED, NO_BONUS, SKIP_ADJUST, NGRAM, RESID_FALLBACK, OFFSET_DEVICE = 1, 2, 4, 8, 16, 32
LANE_OFFSET_DEVICE = 1 << 30

_vp, _i, _i64, _u64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t

_SIGS = {
    "specdec_version": (C.c_int, []),
    "specdec_error_string": (C.c_char_p, [_i]),
    "specdec_workspace_bytes": (_sz, [_i64]),
    "specdec_verify_workspace_bytes": (_sz, [_i, _i, _i]),
    "specdec_sample_rows_workspace_bytes": (_sz, [_i64, _i]),
    "specdec_verify": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _u64, _u64, _i64, _i, _i, _i, _i64, _i64, _i64, _i64,
                            _f, _i, _f, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "specdec_set_profile_events": (_i, [_vp, _vp, _vp]),
    "specdec_debug_stats": (_i, [_vp, _i]),
    "specdec_debug_timeline": (_i, [_vp, _i, _i, _i, _vp, _i]),
    "specdec_set_option": (_i, [C.c_char_p, _i]),
    "specdec_process_probs": (_i, [_vp, _i, _i64, _i, _i64, _f, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "specdec_sample_rows": (_i, [_vp, _i, _i64, _i, _i64, _f, _i, _f, _i, _vp, _u64, _u64, _i64, _i, _vp, _vp,
                                 _vp, _sz, _vp]),
    "specdec_sample_probs": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp]),
    "specdec_philox_uniform": (_i, [_u64, _u64, _i64, _i, _i, _vp, _vp, _vp]),
    "specdec_prune_kv": (_i, [_vp, _i, _i, _i, _i64, _i64, _i, _vp, _vp, _i, _vp]),
    "specdec_ngram_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "specdec_ngram_destroy": (_i, [_vp]),
    "specdec_ngram_reset": (_i, [_vp, _vp]),
    "specdec_ngram_initialize": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp]),
    "specdec_ngram_update": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _i, _vp]),
    "specdec_ngram_lookup_chain": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp]),
    "specdec_ngram_status": (_i, [_vp, _vp]),
    "specdec_ngram_has_gram": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "specdec_ngram_seed": (_i, [_vp, _u64]),
    "specdec_topk_ids": (_i, [_vp, _i, _i64, _i, _i64, _i, _vp, _vp]),
    "specdec_peer_publish": (_i, [_vp, _i, _vp, _i64, _i, _vp, _i64, _i, _vp, _i, _vp, _vp]),
    "specdec_peer_wait": (_i, [_vp, _i, _i, _vp, _vp]),
    "specdec_batch_writeback": (_i, [_i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i, _vp, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C speculative-decoding_b200/csrc`). "
                "There is no CPU / eager fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        # tuning / test hooks of the library, e.g. SPECDEC_OPTS="no_overlap=0,no_fused_tail=1" (see specdec_set_option)
        for kv in filter(None, os.environ.get("SPECDEC_OPTS", "").split(",")):
            k, v = kv.split("=")
            if _lib.specdec_set_option(k.strip().encode(), int(v)) != 0:
                raise RuntimeError(f"SPECDEC_OPTS: unknown option {k!r}")
    return _lib


def exported_symbols():
    return sorted(_SIGS)


class SpecdecError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().specdec_error_string(rc).decode()
        raise SpecdecError(f"{what} failed: {msg} (code {rc})")
