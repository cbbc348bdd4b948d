"""Oracle-driven replays of the reference's three decode loops (CPU, test infrastructure).

Same control flow as sampling/speculative_decoding.py:23-189, ngram_assisted/ngram_assisted.py:11-164 and
engine/infer_engine.py:149-359, with the per-step arithmetic done by the C oracle and the uniforms
popped from injected streams in the order the reference consumes them.  Used to pin the oracle
against golden outputs of the real reference and as the expected value for the GPU drop-ins."""
import numpy as np

from cases import MODES  # noqa: F401


class Streams:
    def __init__(self, sample_u, accept_u):
        self.s = np.asarray(sample_u, np.float32).reshape(-1); self.si = 0
        self.a = np.asarray(accept_u, np.float32).reshape(-1); self.ai = 0

    def sample(self, n=1):
        o = self.s[self.si:self.si + n]; self.si += n; return o

    def accept(self, n):
        o = self.a[self.ai:self.ai + n]; self.ai += n; return o


def speculative_generate(oracle, prompt, q_table, p_table, mode, *, gamma, max_gen_len, eos=-1, pad=0,
                         skip_sample_adjustment=False, first_target=True, sample_u=None, accept_u=None):
    """q_table / p_table: float32 numpy [L, V] position-indexed logits."""
    st = Streams(sample_u, accept_u)
    L = p_table.shape[0]
    stop = eos if isinstance(eos, (list, tuple)) else [eos]
    prompt_len = len(prompt)
    total_len = min(L, prompt_len + max_gen_len)
    ids = [pad] * total_len
    ids[:prompt_len] = list(prompt)
    cp = prompt_len
    acc, spec = 0.0, 0.0

    def samp(row):
        u = [0.0] if mode["greedy"] else st.sample(1)
        return int(oracle.sample_rows(row[None], u, **mode)[0][0])

    if first_target:
        t = samp(p_table[cp - 1])
        ids[cp] = t; cp += 1
        if t in stop:
            return ids[prompt_len:cp], 0
    while cp < total_len:
        g = min(gamma, total_len - cp - 1)
        for k in range(g):
            ids[cp + k] = samp(q_table[cp + k - 1])
        spec += g
        ua = st.accept(g)
        us = [0.0] if mode["greedy"] else st.sample(1)
        o = oracle.verify(p_table[None, cp - 1:cp + g], q_table[None, cp - 1:cp + g - 1] if g > 0 else np.zeros((1, 0, p_table.shape[1]), np.float32),
                          np.asarray(ids[cp:cp + g], np.int64).reshape(1, g), ua, us,
                          flags=4 if skip_sample_adjustment else 0, stop_tokens=stop, **mode)
        n, x, fs = int(o.n_accepted[0]), int(o.next_token[0]), int(o.first_stop[0])
        acc += n
        if fs >= 0:
            return ids[prompt_len:cp + fs + 1], acc / spec
        for k in range(n, g):
            ids[cp + k] = pad
        ids[cp + n] = x
        cp += n + 1
        if x in stop:
            return ids[prompt_len:cp], acc / spec
    return ids[prompt_len:], acc / spec


def ngram_generate(oracle, storage, prompt, p_table, mode, *, gamma, max_gen_len, filler_top_k=3, eos=-1, pad=0,
                   stop_if_unknown=False, sample_u=None, fallback_tokens=None):
    """storage: oracle.ngram_oracle.NGramOracle."""
    st = Streams(sample_u, [])
    L, V = p_table.shape
    stop = eos if isinstance(eos, (list, tuple)) else [eos]
    prompt_len = len(prompt)
    total_len = min(L, prompt_len + max_gen_len)
    ids = [pad] * total_len
    ids[:prompt_len] = list(prompt)
    cp = prompt_len
    acc, spec, fb_i = 0.0, 0.0, 0
    storage.initialize([ids[:prompt_len]])

    def samp(row):
        u = [0.0] if mode["greedy"] else st.sample(1)
        return int(oracle.sample_rows(row[None], u, **mode)[0][0])

    def topk(row, k):
        probs, _ = oracle.process_probs(row[None], temperature=mode["temperature"], top_k=mode["top_k"], top_p=mode["top_p"])
        import torch
        return torch.from_numpy(probs[0]).topk(k).indices.tolist()

    t = samp(p_table[cp - 1])
    ids[prompt_len] = t; cp += 1
    storage.update([ids[:prompt_len]], [[t]])
    while cp < total_len:
        g = min(gamma, total_len - cp - 1)
        cop = list(ids)
        used = g
        for k in range(g):
            fb = int(fallback_tokens[fb_i % len(fallback_tokens)]); fb_i += 1
            tok, known = storage.next_token(cop[:cp + k], 0, fb)
            cop[cp + k] = tok
            if not known and stop_if_unknown:
                used = k
                break
        g = used
        spec += g
        toks = np.asarray(cop[cp:cp + g], np.int64).reshape(1, g)
        if mode["greedy"]:
            ua, us = np.zeros(g, np.float32), [0.0]
            o = oracle.verify(p_table[None, cp - 1:cp + g], None, toks, ua, us, flags=8, stop_tokens=stop, **mode)
        else:
            ua = st.s[st.si:st.si + g]
            ua = np.concatenate([ua, np.zeros(g - len(ua), np.float32)])
            n0 = int(oracle.verify(p_table[None, cp - 1:cp + g], None, toks, ua, [0.0], flags=8, **mode).n_accepted[0])
            st.si += min(n0 + 1, g)
            us = st.sample(1)
            o = oracle.verify(p_table[None, cp - 1:cp + g], None, toks, ua, us, flags=8, stop_tokens=stop, **mode)
        n, x, fs = int(o.n_accepted[0]), int(o.next_token[0]), int(o.first_stop[0])
        acc += n
        if fs >= 0:
            return cop[prompt_len:cp + fs + 1], acc / spec
        ids[cp:cp + n] = cop[cp:cp + n]
        ids[cp + n] = x
        for i in range(n):
            storage.update([ids[:cp + i]], [[ids[cp + i]]])
            if filler_top_k > 1:
                storage.update([ids[:cp + i]], [topk(p_table[cp - 1 + i], filler_top_k)])
        storage.update([ids[:cp + n]], [[x]])
        if filler_top_k > 1:
            storage.update([ids[:cp + n]], [topk(p_table[cp - 1 + n], filler_top_k)])
        cp += n + 1
        if x in stop:
            return ids[prompt_len:cp], (acc / spec if spec > 0 else 0.0)
    return ids[prompt_len:], (acc / spec if spec > 0 else 0.0)


def batch_generate(oracle, input_ids, q_table, p_table, *, gamma, gen_len, end_tokens=(), sample_u=None, accept_u=None):
    """engine/infer_engine.py:149-359 with T=1 / no processor; q_table, p_table float32 [B, L, V]."""
    st = Streams(sample_u, accept_u)
    mode = dict(temperature=1.0, top_k=0, top_p=1.0, greedy=False)
    FL = 1 | 2 | 16
    B, P = input_ids.shape
    G = gen_len
    gen = np.zeros((B, G), np.int64)
    fin = np.zeros(B, bool)
    ngen = np.zeros(B, np.int64); nacc = np.zeros(B, np.int64)
    step = 0
    while step < G:
        if fin.all():
            break
        g = min(gamma, G - step)
        dtok = np.zeros((B, g), np.int64)
        for k in range(g):
            rows = q_table[:, P + step + k]                           # the drafter has consumed P+step+k tokens
            tok = oracle.sample_rows(rows, st.sample(B), **mode)[0]
            act = ~fin
            dtok[act, k] = tok[act]; gen[act, step + k] = tok[act]; ngen[act] += 1
        for b in range(B):
            if fin[b]:
                continue
            tl = p_table[b:b + 1, P + step - 1:P + step + g - 1]      # logits[:, -(g+1):-1]
            dl = q_table[b:b + 1, P + step:P + step + g]
            ua = st.a[st.ai:st.ai + g]
            ua = np.concatenate([ua, np.zeros(g - len(ua), np.float32)])
            pad_t = np.concatenate([tl, tl[:, :1]], 1)                # oracle wants g+1 rows; bonus unused
            o0 = oracle.verify(pad_t, dl, dtok[b:b + 1], ua, [0.0], flags=FL, stop_tokens=end_tokens, **mode)
            n, fs = int(o0.n_accepted[0]), int(o0.first_stop[0])
            used = (fs + 1) if fs >= 0 else min(n + 1, g)
            st.ai += used
            if fs >= 0:
                acc = fs + 1; fin[b] = True
            else:
                acc = n
                if n < g:
                    us = st.sample(1)
                    x = int(oracle.verify(pad_t, dl, dtok[b:b + 1], ua, us, flags=FL, stop_tokens=end_tokens, **mode).next_token[0])
                    gen[b, step + n] = x
                    if x in end_tokens:
                        fin[b] = True
            nacc[b] += acc
            if acc < g and step + acc + 1 < step + g:
                gen[b, step + acc + 1:step + g] = 0
        step += g
    outs = []
    for b in range(B):
        nz = np.nonzero(gen[b])[0]
        outs.append(list(input_ids[b]) + (gen[b, :nz[-1] + 1].tolist() if nz.size else []))
    return outs, [(nacc[b] / ngen[b]) if ngen[b] > 0 else 0.0 for b in range(B)]
