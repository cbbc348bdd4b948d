"""Drop-in batch_speculative_generate (engine/infer_engine.py:149-359): same signature, same `ctx`
attribute bag (.drafter .target .gamma .gen_len .end_tokens .target_device), same return value
(List[Tensor] prompt+generated, List[float] acceptance rates) and the same token-level semantics,
including the reference's quirks (no bonus token, `step += gamma` regardless of the accepted
count, zero-filled tail, only trailing zeros trimmed).

What changes is the hot loop: the reference materialises q_probs_full/p_probs_full [B,gamma,V] fp32
(:221,:276) and then runs B*gamma Python iterations with >= 3 .item() syncs each (:280-336).  Here the
drafter step is one specdec::sample_rows launch, the whole accept / residual-resample block is ONE
specdec::verify launch (flags ACCEPT_BATCHED|NO_BONUS|RESID_FALLBACK) on the drafter's *logits*, and
the write-back (accepted counts, corrected token, zeroed tail, finished flags) is ONE
specdec_batch_writeback launch on device-resident state.  The Philox offset lives on the device and the
"all finished" exit is polled from a pinned flag one step late, so no step waits for the host.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops
from .uniforms import PhiloxUniforms, default_uniforms  # noqa: F401

_FLAGS = ops.L.ACCEPT_BATCHED | ops.L.NO_BONUS | ops.L.RESID_FALLBACK


@torch.no_grad()
def batch_speculative_generate(ctx, input_ids: torch.Tensor, attention_mask: torch.Tensor, batch_size: int,
                               first_token_callback=None, uniforms=None, seq_id0: int = 0
                               ) -> Tuple[List[torch.Tensor], List[float]]:
    un = uniforms if uniforms is not None else default_uniforms()
    device = input_ids.device
    target_device = getattr(ctx, "target_device", device)
    B, G = batch_size, ctx.gen_len
    generated = torch.zeros(B, G, device=device, dtype=torch.long)
    finished = torch.zeros(B, dtype=torch.bool, device=device)
    n_gen = torch.zeros(B, dtype=torch.long, device=device)
    n_acc = torch.zeros(B, dtype=torch.long, device=device)
    end_tokens = torch.as_tensor(list(ctx.end_tokens), dtype=torch.long, device=device)

    out0 = ctx.drafter(input_ids, attention_mask=attention_mask, use_cache=True)
    drafter_past = out0.past_key_values

    # ---- device-resident loop state: nothing below reads a device value on the host per step.
    #  * Philox offset: one int64 word on the device, bumped by one tiny in-stream add per step; every kernel of the
    #    step reads it when it RUNS (SPECDEC_OFFSET_DEVICE), lanes keep the calls of one step apart.
    #  * finished / n_acc / generated are updated by ONE specdec_batch_writeback launch per step.
    #  * the "all finished" early exit (:211) is polled one step late from a pinned flag (non-blocking copy + event):
    #    an extra step on finished sequences changes nothing (they are masked out), so results are identical.
    dev_off = None
    if not un.injected:
        dev_off = torch.full((1,), un.offset, dtype=torch.int64, device=device)
    n_active = torch.zeros(1, dtype=torch.int32, device=device)
    n_active_host = torch.ones(1, dtype=torch.int32).pin_memory() if device.type == "cuda" else None
    poll_event = None
    steps_run = 0

    step = 0
    while step < G:
        if un.injected:
            if bool(finished.all()):
                break
        elif poll_event is not None and poll_event.query() and int(n_active_host[0]) == 0:
            break
        g = min(ctx.gamma, G - step)
        active = ~finished
        draft_tokens = torch.zeros(B, g, device=device, dtype=torch.long)
        draft_logits = None
        # ---- drafter: g sequential 1-token forwards, fused softmax+sample per step (:224-263)
        for k in range(g):
            if k == 0:
                prev = generated[:, step - 1] if step > 0 else input_ids[:, -1]
            else:
                prev = generated[:, step + k - 1]
            out = ctx.drafter(prev.unsqueeze(1), past_key_values=drafter_past, use_cache=True)
            drafter_past = out.past_key_values
            logits = out.logits[:, -1, :]
            if draft_logits is None:
                draft_logits = torch.empty(B, g, logits.shape[-1], dtype=logits.dtype, device=device)
            draft_logits[:, k] = logits
            if un.injected:
                tok, _ = ops.sample_rows(logits, un.sample(B))
            else:
                tok, _ = ops.sample_rows(logits, None, seed=un.seed, offset=dev_off, seq_id0=seq_id0, lane_id=k)
            tok = tok.to(device)
            draft_tokens[:, k] = torch.where(active, tok, draft_tokens[:, k])
            generated[:, step + k] = torch.where(active, tok, generated[:, step + k])
            n_gen += active.long()
            if first_token_callback is not None and k == 0 and step == 0:
                for idx in torch.where(active)[0]:
                    first_token_callback(idx.item())
        # ---- target: full re-forward like the reference (:270-275), then ONE fused verify
        verify_ids = torch.cat([input_ids, generated[:, :step + g]], dim=1).to(target_device)
        t_logits = ctx.target(verify_ids).logits[:, -(g + 1):-1, :]
        if t_logits.device != device:
            t_logits = t_logits.to(device)
        if un.injected:
            res = _verify_injected(un, t_logits, draft_logits, draft_tokens, active, end_tokens)
        else:
            res = ops.fused_verify(t_logits, draft_logits, draft_tokens, None, None, seed=un.seed,
                                   offset=dev_off, seq_id0=seq_id0, flags=_FLAGS, stop_tokens=end_tokens)
        # ---- accepted counts, corrected token, zeroed tail, finished flags (:300-336): one launch, no read-back
        ops.batch_writeback(res, generated, step, g, finished, n_acc, end_tokens, n_active)
        if not un.injected:
            dev_off.add_(1)
            steps_run += 1
            n_active_host.copy_(n_active, non_blocking=True)
            poll_event = torch.cuda.Event()
            poll_event.record()
        step += g
    if not un.injected:
        un.offset += steps_run

    outs: List[torch.Tensor] = []
    rates: List[float] = []
    gen_cpu, ng, na = generated.cpu(), n_gen.tolist(), n_acc.tolist()
    for i in range(B):
        nz = torch.nonzero(gen_cpu[i], as_tuple=True)[0]
        final = generated[i, :int(nz[-1]) + 1] if nz.numel() > 0 else torch.tensor([], dtype=torch.long, device=device)
        outs.append(torch.cat([input_ids[i], final]))
        rates.append((na[i] / ng[i]) if ng[i] > 0 else 0.0)
    return outs, rates


def _verify_injected(un, t_logits, draft_logits, draft_tokens, active, end_tokens):
    """Test-only: replay the reference's lazily drawn torch.rand(1) / torch.multinomial streams
    (engine/infer_engine.py:305,321,325): sequence by sequence, one uniform per decision made."""
    B, g, _ = draft_logits.shape
    dev = draft_logits.device
    outs = []
    for b in range(B):
        if not bool(active[b]):
            outs.append(None)
            continue
        ua = un.peek_accept(g)
        r0 = ops.fused_verify(t_logits[b:b + 1], draft_logits[b:b + 1], draft_tokens[b:b + 1], ua,
                              torch.zeros(1, device=dev), flags=_FLAGS, stop_tokens=end_tokens)
        n, fs = int(r0.n_accepted[0]), int(r0.first_stop[0])
        used = (fs + 1) if fs >= 0 else min(n + 1, g)
        un.skip_accept(used)
        us = un.sample(1) if (fs < 0 and n < g) else torch.zeros(1, device=dev)
        outs.append(ops.fused_verify(t_logits[b:b + 1], draft_logits[b:b + 1], draft_tokens[b:b + 1], ua, us,
                                     flags=_FLAGS, stop_tokens=end_tokens))

    class R:
        pass
    r = R()
    z32 = torch.zeros(1, dtype=torch.int32, device=dev)
    r.n_accepted = torch.cat([o.n_accepted if o else z32 for o in outs])
    r.first_stop = torch.cat([o.first_stop if o else z32 - 1 for o in outs])
    r.next_token = torch.cat([o.next_token if o else torch.zeros(1, dtype=torch.int64, device=dev) for o in outs])
    return r
