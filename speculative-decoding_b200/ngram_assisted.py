"""Drop-in ngram_assisted_speculative_generate (ngram_assisted/ngram_assisted.py:11-164): n-gram
drafts (one chained-lookup launch), exact-match accept `draft == sample(p_i)` (:114-119), next token
from p[n] / bonus row (:132-141) -- all inside one specdec::verify call with SPECDEC_NGRAM -- then
table updates with the top-k filler (:149-155)."""
from __future__ import annotations

from typing import List

import torch
from torch.nn import Module

from . import ops
from .caching import prune_cache
from .logits_processor import LogitsProcessor, GreedyProcessor
from .ngram_storage import INgramStorage
from .uniforms import PhiloxUniforms, default_uniforms  # noqa: F401


@torch.no_grad()
def ngram_assisted_speculative_generate(
    inputs: List[int],
    ngramstorage: INgramStorage,
    target: Module,
    tokenizer=None,
    gamma: int = 5,
    filler_top_k: int = 3,
    logits_processor: LogitsProcessor = None,
    max_gen_len: int = 40,
    eos_tokens_id: int | List[int] = 1,
    pad_token_id: int = 0,
    use_cache: bool = False,
    first_target: bool = True,
    stop_if_unknown: bool = False,
    debug: bool = False,
    uniforms=None,
    fallback_tokens=None,
):
    if logits_processor is None:
        logits_processor = GreedyProcessor()
    fp = logits_processor.fused_params()
    un = uniforms if uniforms is not None else default_uniforms()
    dev = target.device
    target_cache = None
    list_tokens_id = eos_tokens_id if isinstance(eos_tokens_id, list) else [eos_tokens_id]
    stop_tokens = torch.tensor(list_tokens_id, dtype=torch.long, device=dev)
    stop_set = set(int(t) for t in list_tokens_id)
    drafts_accepted, drafts_speculated = .0, .0
    fb_i = 0

    prompt_len = len(inputs)
    total_len = min(target.config.max_position_embeddings, prompt_len + max_gen_len)
    input_ids = torch.full((1, total_len), pad_token_id, dtype=torch.long, device=dev)
    input_ids[0, :prompt_len] = torch.tensor(inputs, dtype=torch.long, device=dev)
    current_position = prompt_len

    ngramstorage.initialize(input_ids[..., :prompt_len])

    def _sample(row):
        if un.injected:
            u = None if fp["greedy"] else un.sample(1)
            return ops.sample_rows(row, u, **fp)[0]
        return ops.sample_rows(row, None, seed=un.seed, offset=un.next_offset(), **fp)[0]

    if first_target:
        Mp = target(input_ids=input_ids[..., :current_position], past_key_values=target_cache, use_cache=use_cache)
        target_cache = Mp.past_key_values
        t = _sample(Mp.logits[..., -1, :])
        input_ids[0, prompt_len] = t[0]
        current_position += 1
        ngramstorage.update(input_ids[..., :prompt_len], t.reshape(1, 1))

    while current_position < total_len:
        corrected_gamma = min(gamma, total_len - current_position - 1)
        g = corrected_gamma
        fb = None
        if fallback_tokens is not None and g > 0:
            fb = torch.tensor([[int(fallback_tokens[(fb_i + t) % len(fallback_tokens)]) for t in range(g)]], device=dev)
        copied = input_ids.clone()
        if g > 0:
            drafts, known = ngramstorage.lookup_chain(input_ids[..., :current_position], g, fallback=fb)
            kn = known[0].tolist()
            used = g
            if stop_if_unknown:
                for k in range(g):
                    if not kn[k]:
                        used = k
                        break
            # the reference draws one fallback per next_token() call, also for the call that stops
            fb_i += min(g, used + 1) if stop_if_unknown else g
            g = used
            copied[0, current_position:current_position + g] = drafts[0, :g]
        drafts_speculated += g

        Mp = target(input_ids=copied[..., :current_position + g], past_key_values=target_cache, use_cache=use_cache)
        target_cache = Mp.past_key_values
        tl = Mp.logits[:, current_position - 1:current_position + g, :]
        toks = copied[:, current_position:current_position + g]
        if un.injected:
            # the reference draws sample() lazily until the first mismatch (ngram_assisted.py:114-119),
            # so replaying its uniform stream needs n before the final sample's uniform is known
            if fp["greedy"]:
                ua = torch.zeros(g, device=dev) if g > 0 else None
                us = torch.zeros(1, device=dev)
            else:
                ua = un.peek_sample(g) if g > 0 else None
                n0 = int(ops.fused_verify(tl, None, toks, ua, torch.zeros(1, device=dev), flags=ops.L.NGRAM, **fp).n_accepted[0])
                un.skip_sample(min(n0 + 1, g))
                us = un.sample(1)
            res = ops.fused_verify(tl, None, toks, ua, us, flags=ops.L.NGRAM, stop_tokens=stop_tokens, **fp)
        else:
            res = ops.fused_verify(tl, None, toks, None, None, seed=un.seed, offset=un.next_offset(),
                                   flags=ops.L.NGRAM, stop_tokens=stop_tokens, **fp)
        # filler ids of every position of the step in one launch, issued BEFORE the (single) host read-back so that it
        # runs while the host waits: ids of the filler_top_k largest logits per row (ngram_assisted.py:149-155)
        fill = ops.topk_ids(tl[0], filler_top_k) if filler_top_k > 1 else None
        hn, hx, hf = res.host()
        n, x, fs = hn[0], hx[0], hf[0]
        drafts_accepted += n
        if fs >= 0:
            return copied[0, prompt_len:current_position + fs + 1].tolist(), drafts_accepted / drafts_speculated
        if n < g and use_cache:
            target_cache = prune_cache(target_cache, g - n + 1)
        input_ids[0, current_position:current_position + n] = copied[0, current_position:current_position + n]
        input_ids[0, current_position + n] = x

        # update the ngram model (ngram_assisted.py:149-155)
        if hasattr(ngramstorage, "update_chain"):
            # device tables: the 2 (n + 1) updates of the step, in the reference's order, as ONE launch
            ngramstorage.update_chain(input_ids[0], current_position, input_ids[0, current_position:current_position + n + 1],
                                      fill[:n + 1] if filler_top_k > 1 else None)
        else:  # any other INgramStorage: one call per update, as the reference does
            for i in range(n):
                ngramstorage.update(input_ids[..., :current_position + i], input_ids[..., current_position + i].reshape(1, 1))
                if filler_top_k > 1:
                    ngramstorage.update(input_ids[..., :current_position + i], fill[i].reshape(1, -1))
            ngramstorage.update(input_ids[..., :current_position + n], torch.tensor([[x]], device=dev))
            if filler_top_k > 1:
                ngramstorage.update(input_ids[..., :current_position + n], fill[n].reshape(1, -1))

        current_position += n + 1
        if x in stop_set:
            return input_ids[0, prompt_len:current_position].tolist(), (drafts_accepted / drafts_speculated
                                                                        if drafts_speculated > 0 else 0.0)
    return input_ids[0, prompt_len:].tolist(), (drafts_accepted / drafts_speculated if drafts_speculated > 0 else 0.0)
