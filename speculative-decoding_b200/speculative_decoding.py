"""Drop-in speculative_generate (sampling/speculative_decoding.py:23-189 of the reference): same
signature, same control flow and return values; the per-step arithmetic -- drafter
processor+sample (:120-124), target processor (:136), accept loop (:139-145), stop scan
(:150-155), bonus / residual resample (:158-171) -- is one specdec::sample_rows call per draft and
ONE specdec::verify call per step, with a single host read-back of the packed result.
Nothing V-sized is ever materialised (the reference allocates q[1,gamma,V] fp32 every step, :107).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
from torch.nn import Module

from . import ops
from .caching import prune_cache
from .logits_processor import LogitsProcessor, GreedyProcessor
from .uniforms import PhiloxUniforms, default_uniforms  # noqa: F401


def max_fn(x: torch.Tensor) -> torch.Tensor:
    """norm(max(0, x)) -- kept for API parity (sampling/speculative_decoding.py:10-19); the fused
    verify op computes it inside the kernel."""
    x_max = torch.where(x > 0, x, torch.zeros_like(x))
    return x_max / torch.sum(x_max, dim=-1, keepdim=True)


@torch.no_grad()
def speculative_generate(
    inputs: List[int],
    drafter: Module,
    target: Module,
    tokenizer=None,
    gamma: int = 5,
    logits_processor: LogitsProcessor = None,
    max_gen_len: int = 40,
    eos_tokens_id: int | List[int] = 1,
    pad_token_id: int = 0,
    use_cache: bool = False,
    skip_sample_adjustment: bool = False,
    first_target: bool = True,
    debug: bool = False,
    uniforms=None,
) -> Tuple[List[int], float]:
    """Generate text with speculative decoding (https://arxiv.org/pdf/2211.17192.pdf), batch size 1.

    `uniforms`: optional PhiloxUniforms (default, seeded from torch.initial_seed()) or
    InjectedUniforms (tests).  All other arguments as in the reference."""
    if logits_processor is None:
        logits_processor = GreedyProcessor()
    fp = logits_processor.fused_params()
    un = uniforms if uniforms is not None else default_uniforms()
    dev = target.device
    drafter_cache, target_cache = None, None

    list_tokens_id = eos_tokens_id if isinstance(eos_tokens_id, list) else [eos_tokens_id]
    stop_tokens = torch.tensor(list_tokens_id, dtype=torch.long, device=dev)
    stop_set = set(int(t) for t in list_tokens_id)

    drafts_accepted, drafts_speculated = .0, .0

    prompt_len = len(inputs)
    cfg = target.config
    max_seq_length = cfg.max_position_embeddings if hasattr(cfg, 'max_position_embeddings') else (
        cfg.max_context_length if hasattr(cfg, 'max_context_length') else 1024)
    total_len = min(max_seq_length, prompt_len + max_gen_len)
    input_ids = torch.full((1, total_len), pad_token_id, dtype=torch.long, device=dev)
    input_ids[0, :prompt_len] = torch.tensor(inputs, dtype=torch.long, device=dev)
    current_position = prompt_len

    def _sample(logits_row, lane):
        """processor + sample() on one [1,V] row -> token tensor [1]"""
        if un.injected:
            u = None if fp["greedy"] else un.sample(1)
            tok, _ = ops.sample_rows(logits_row, u, **fp)
        else:
            tok, _ = ops.sample_rows(logits_row, None, seed=un.seed, offset=un.next_offset(), lane_id=lane, **fp)
        return tok

    if first_target:
        Mp = target(input_ids=input_ids[..., :current_position], past_key_values=target_cache, use_cache=use_cache)
        target_cache = Mp.past_key_values
        t = _sample(Mp.logits[..., -1, :], 0)
        input_ids[0, current_position] = t[0]
        current_position += 1
        if int(t[0]) in stop_set:
            return input_ids[0, prompt_len:current_position].tolist(), 0

    while current_position < total_len:
        corrected_gamma = min(gamma, total_len - current_position - 1)
        g = corrected_gamma
        draft_logits = []
        input_ids = input_ids.to(drafter.device)
        for k in range(g):
            Mq = drafter(input_ids=input_ids[..., :current_position + k], past_key_values=drafter_cache,
                         use_cache=use_cache)
            drafter_cache = Mq.past_key_values
            row = Mq.logits[..., -1, :]
            draft_logits.append(row.to(dev))
            xi = _sample(row, k)
            input_ids[0, current_position + k] = xi[0]
        drafts_speculated += g
        input_ids = input_ids.to(dev)

        Mp = target(input_ids=input_ids[..., :current_position + g], past_key_values=target_cache, use_cache=use_cache)
        target_cache = Mp.past_key_values
        tl = Mp.logits[:, current_position - 1:current_position + g, :]  # gamma rows + bonus row
        if g > 0:
            dl = torch.stack([r.reshape(-1) for r in draft_logits], 0).unsqueeze(0)
        else:
            dl = None
        toks = input_ids[:, current_position:current_position + g]
        flags = ops.L.SKIP_ADJUST if skip_sample_adjustment else 0
        if un.injected:
            ua = un.accept(g) if g > 0 else None
            us = un.sample(1) if not fp["greedy"] else torch.zeros(1, device=dev)
            res = ops.fused_verify(tl, dl, toks, ua, us, flags=flags, stop_tokens=stop_tokens, **fp)
        else:
            res = ops.fused_verify(tl, dl, toks, None, None, seed=un.seed, offset=un.next_offset(), flags=flags,
                                   stop_tokens=stop_tokens, **fp)
        # one host sync per step
        hn, hx, hf = res.host()
        n, x, fs = hn[0], hx[0], hf[0]
        drafts_accepted += n

        if fs >= 0:  # an accepted draft is a stop token (sampling/speculative_decoding.py:150-155)
            return input_ids[0, prompt_len:current_position + fs + 1].tolist(), drafts_accepted / drafts_speculated

        if n < g and use_cache:
            drafter_cache = prune_cache(drafter_cache, g - n)
            target_cache = prune_cache(target_cache, g - n + 1)

        input_ids[0, current_position + n:current_position + g] = pad_token_id
        input_ids[0, current_position + n] = x
        current_position += n + 1

        if x in stop_set:
            return input_ids[0, prompt_len:current_position].tolist(), drafts_accepted / drafts_speculated

    return input_ids[0, prompt_len:].tolist(), drafts_accepted / drafts_speculated
