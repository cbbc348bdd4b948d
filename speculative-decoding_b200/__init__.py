"""specdec_b200 -- B200-native (sm_100a) speculative-sampling verify path with the API of
dadiaokua/speculative-decoding: LogitsProcessor family, speculative_generate,
ngram_assisted_speculative_generate, batch_speculative_generate, prune_cache, NGramStorage.

The compute lives in libspecdec_b200.so (hand-written CUDA behind a C ABI, include/specdec_b200.h),
registered as torch.library custom ops (`torch.ops.specdec.*`).  No Triton, no multi-backend
dispatch, no CPU fallback: ops raise if the library is missing or tensors are not on CUDA.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401  (registers torch.ops.specdec.*)
from .ops import fused_verify, process_probs, sample_rows, sample_probs, philox_uniform, prune_kv, VerifyResult, GraphedVerify, topk_ids, batch_writeback  # noqa: F401
from .uniforms import PhiloxUniforms, InjectedUniforms, default_uniforms  # noqa: F401
from .logits_processor import (LogitsProcessor, GreedyProcessor, MultinomialProcessor, TopKProcessor,  # noqa: F401
                               NucleusProcessor, TopKNucleusProcessor)
from .caching import prune_cache, prune_tuple_cache, prune_dynamic_cache, StaticKVCache  # noqa: F401
from .speculative_decoding import speculative_generate, max_fn  # noqa: F401
from .ngram_storage import INgramStorage, NGramStorage, OneLevelNGramStorage  # noqa: F401
from .ngram_assisted import ngram_assisted_speculative_generate  # noqa: F401
from .infer_engine import batch_speculative_generate  # noqa: F401
from . import dist  # noqa: F401

__version__ = "0.1.0"
