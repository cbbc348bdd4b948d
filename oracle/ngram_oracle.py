"""CPU oracle for the n-gram tables (TEST INFRASTRUCTURE ONLY): a plain-Python restatement of
NGramStorage / OneLevelNGramStorage (ngram_assisted/ngram_storage.py:73-249) with the two
extensions the device tables have: one logical table per table id, and deterministic fallback
tokens instead of torch.randint (:84,:165).  Pinned against the reference classes by
tests/test_oracle_vs_reference.py / tests/golden/ngram_*.npz."""
from __future__ import annotations


class NGramOracle:
    def __init__(self, n, vocab_size, one_level=False):
        assert n > 1
        self.n, self.vocab_size, self.one_level = n, vocab_size, one_level
        self.reset()

    def reset(self):
        self.counts = {}  # table -> j -> gram -> {token: count}
        self.best = {}    # table -> j -> gram -> token

    def _observe(self, tab, j, gram, tokens):
        c = self.counts.setdefault(tab, {}).setdefault(j, {})
        b = self.best.setdefault(tab, {}).setdefault(j, {})
        if gram not in c:
            c[gram] = {}
        if gram not in b:
            b[gram] = tokens[0]
        for t in tokens:
            if t not in c[gram]:
                c[gram][t] = 1                      # ngram_storage.py:215-216 (no arg-max update)
            else:
                c[gram][t] += 1
                if c[gram][t] > c[gram][b[gram]]:   # strict >: incumbent kept on ties (:220)
                    b[gram] = t

    def update(self, seqs, next_tokens, table_ids=None):
        for i, seq in enumerate(seqs):
            seq = list(seq)
            tab = 0 if table_ids is None else int(table_ids[i])
            toks = [int(t) for t in next_tokens[i]]
            if self.one_level:
                if len(seq) < self.n:               # ngram_storage.py:110
                    continue
                self._observe(tab, self.n - 1, tuple(seq[-(self.n - 1):]), toks)
            else:
                if len(seq) < 1:
                    continue
                for j in range(min(self.n - 1, len(seq)), 1, -1):   # :200
                    self._observe(tab, j, tuple(seq[-j:]), toks)

    def initialize(self, seqs, table_ids=None):
        for i, seq in enumerate(seqs):
            seq = list(seq)
            tab = 0 if table_ids is None else int(table_ids[i])
            if self.one_level:
                for k in range(len(seq) - self.n + 1):             # :132-146
                    self._observe(tab, self.n - 1, tuple(seq[k:k + self.n - 1]), [int(seq[k + self.n - 1])])
            else:
                for k in range(len(seq)):                          # :225-245
                    for j in range(min(self.n - 1, k), 1, -1):
                        self._observe(tab, j, tuple(seq[k - j:k]), [int(seq[k])])

    def next_token(self, seq, tab=0, fallback=0):
        seq = list(seq)
        b = self.best.get(tab, {})
        if self.one_level:
            if len(seq) >= self.n - 1:
                g = tuple(seq[-(self.n - 1):]) if self.n > 1 else ()
                if g in b.get(self.n - 1, {}):
                    return b[self.n - 1][g], True
            return fallback, False
        for j in range(min(self.n - 1, len(seq)), 1, -1):          # :171-177
            g = tuple(seq[-j:])
            if g in b.get(j, {}):
                return b[j][g], True
        return fallback, False

    def has_gram(self, ngram, tab=0):
        """ngram_storage.py:98-106 (one level) / :181-193: the context is the LAST j tokens of `ngram`, its final token
        included, and the question is whether that final token was ever counted after it."""
        ngram = [int(t) for t in ngram]
        c = self.counts.get(tab, {})
        if self.one_level:
            if len(ngram) < self.n:
                return False
            g = tuple(ngram[-(self.n - 1):])
            return g in c.get(self.n - 1, {}) and ngram[-1] in c[self.n - 1][g]
        if len(ngram) < 1:
            return False
        for j in range(min(self.n - 1, len(ngram)), 1, -1):
            g = tuple(ngram[-j:])
            if g in c.get(j, {}) and ngram[-1] in c[j][g]:
                return True
        return False

    def lookup_chain(self, seqs, gamma, table_ids=None, fallback=None):
        drafts, known = [], []
        for i, seq in enumerate(seqs):
            seq = list(seq)
            tab = 0 if table_ids is None else int(table_ids[i])
            d, k = [], []
            for s in range(gamma):
                fb = 0 if fallback is None else int(fallback[i][s])
                t, kn = self.next_token(seq + d, tab, fb)
                d.append(int(t)); k.append(bool(kn))
            drafts.append(d); known.append(k)
        return drafts, known
