/*
 * specdec_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the speculative-sampling verify path of
 * dadiaokua/speculative-decoding, used ONLY by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs as the checker for the
 * CUDA path.  Nothing under speculative-decoding_b200/ may import or link it.
 *
 * Parity status: the reference ships no tests / golden vectors ("parity
 * unpinned by the reference's own tests").  This oracle is
 * pinned instead against outputs of the reference itself, generated in the
 * build container by tests/golden/make_golden.py (which imports
 * /root/reference) and committed as tests/golden/ (npz files).
 *
 * What it restates (reference file:line):
 *   utils/logits_processor.py:13-15   softmax(_process(logits)/T)       -> row_stats()/row_prob()
 *   utils/logits_processor.py:35-36   greedy sample = argmax, first idx  -> sample_p_row(greedy)
 *   utils/logits_processor.py:59-63   top-k: keep logits >= k-th largest -> row_stats() top-k part
 *   utils/logits_processor.py:73-81   nucleus (T=1 cumsum, shifted mask) -> row_stats() nucleus part
 *   utils/logits_processor.py:92-103  top-k then nucleus                 -> both
 *   sampling/speculative_decoding.py:10-19   max_fn = norm(max(0,x))    -> sample_residual()
 *   sampling/speculative_decoding.py:139-145 r > p/q first rejection     -> oracle_verify() rule 0
 *   sampling/speculative_decoding.py:150-155 stop-token scan             -> first_stop
 *   sampling/speculative_decoding.py:158-171 bonus / residual / skip     -> oracle_verify()
 *   engine/infer_engine.py:297-326    min(1,p/q), strict <, fallback     -> rule 1 / RESID_FALLBACK
 *   ngram_assisted/ngram_assisted.py:114-141 accept iff draft==sample(p) -> NGRAM flag
 *
 * Canonical arithmetic.  torch's softmax / cumsum have implementation-defined
 * summation order and libm exp, so "bit-exact emitted tokens" needs an exactly
 * specified arithmetic that a GPU can reproduce.  The spec:
 *   c      = (float)(log2(e) / (double)T)
 *   t_j    = fmaf(z_j, c, -(m*c))            m = row max (fp32), one rounding
 *   e_j    = cexp2(t_j)                       degree-5 polynomial, IEEE ops only
 *   Sfix   = sum_j (uint64)(e_j * 2^40)       integer => order independent
 *   S32    = (float)Sfix * 2^-40 ; inv = 1.0f/S32 ; P_j = e_j * inv
 *   accept : !(u > P_tok / Q_tok)            IEEE fp32 division
 *   resid  : r_j = max(0, P_j - Q_j) ; Rfix = sum (uint64)(r_j * 2^60)
 *   invCDF : target = (Rfix * floor(u*2^24)) >> 24 ; first j with cum > target
 * Every step is IEEE-754 fp32 (+, *, fma, /) or integer arithmetic, so the CUDA
 * kernels reproduce it bit for bit regardless of launch geometry.
 *
 * Build: see oracle/Makefile (gcc -O2 -mfma -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* flags (mirror include/specdec_b200.h, restated here on purpose) */
#define F_ACCEPT_BATCHED 1   /* engine/infer_engine.py:303-305 rule         */
#define F_NO_BONUS 2         /* engine/infer_engine.py: no bonus token      */
#define F_SKIP_ADJUST 4      /* skip_sample_adjustment                      */
#define F_NGRAM 8            /* ngram_assisted.py: accept iff tok==sample   */
#define F_RESID_FALLBACK 16  /* infer_engine.py:319 denom<=1e-12 -> p       */

static const float C1 = 0x1.62e42ap-1f, C2 = 0x1.ebf9bcp-3f, C3 = 0x1.c6b752p-5f,
                   C4 = 0x1.3cea88p-7f, C5 = 0x1.5bba14p-10f;
static const double LOG2E = 1.4426950408889634;

typedef unsigned __int128 u128;

static inline float cexp2(float t) {
  t = fmaxf(t, -125.0f);
  t = fminf(t, 126.0f);
  float r = t + 12582912.0f;
  int32_t ri;
  memcpy(&ri, &r, 4);
  int32_t i = ri - 0x4B400000;
  float fi = r - 12582912.0f;
  float f = t - fi;
  float p = C5;
  p = fmaf(p, f, C4);
  p = fmaf(p, f, C3);
  p = fmaf(p, f, C2);
  p = fmaf(p, f, C1);
  p = fmaf(p, f, 1.0f);
  int32_t pb;
  memcpy(&pb, &p, 4);
  pb += (int32_t)((uint32_t)i << 23);
  float out;
  memcpy(&out, &pb, 4);
  return out;
}
static inline uint64_t fix40(float x) { return (uint64_t)(x * 1099511627776.0f); }
static inline uint64_t fix60(float x) { return (uint64_t)(x * 1152921504606846976.0f); }
static inline uint32_t u24_of(float u) {
  if (!(u > 0.0f)) return 0;
  if (u >= 1.0f) return 16777215u;
  return (uint32_t)(u * 16777216.0f);
}
static inline uint64_t scale_u24(uint64_t S, uint32_t u24) { return (uint64_t)(((u128)S * u24) >> 24); }

typedef struct {
  float m;        /* row max                                     */
  float mc;       /* m * c                                       */
  float inv;      /* 1 / S32                                     */
  float S32;
  uint64_t Sfix;  /* sum fix40(e_j) over kept                    */
  float kth;      /* top-k threshold (-inf when off)             */
  float cut;      /* nucleus cut value (-inf when off)           */
  int64_t jcut;   /* ties at cut kept iff index <= jcut          */
  int64_t n_kept;
} RowStats;

static inline int row_kept(const RowStats* s, float z, int64_t j) {
  if (z < s->kth) return 0;
  if (z > s->cut) return 1;
  return z == s->cut && j <= s->jcut;
}

typedef struct { float z; int32_t j; } ZI;
static int cmp_desc(const void* a, const void* b) {
  const ZI* x = (const ZI*)a; const ZI* y = (const ZI*)b;
  if (x->z > y->z) return -1;
  if (x->z < y->z) return 1;
  return (x->j > y->j) - (x->j < y->j);
}

/* utils/logits_processor.py:13-15,59-63,73-81,92-103 in canonical arithmetic */
static void row_stats(const float* z, int64_t V, float c, int top_k, float top_p, RowStats* s) {
  float m = -INFINITY;
  for (int64_t j = 0; j < V; ++j) m = fmaxf(m, z[j]);
  s->m = m; s->mc = m * c;
  s->kth = -INFINITY; s->cut = -INFINITY; s->jcut = V;
  int use_k = top_k > 0 && top_k < V;
  int use_p = top_p > 0.0f && top_p < 1.0f;
  if (use_k || use_p) {
    ZI* a = (ZI*)malloc(sizeof(ZI) * (size_t)V);
    for (int64_t j = 0; j < V; ++j) { a[j].z = z[j]; a[j].j = (int32_t)j; }
    qsort(a, (size_t)V, sizeof(ZI), cmp_desc);
    int64_t nk = V;
    if (use_k) {                       /* logits < topk(...)[-1] removed: ties kept */
      s->kth = a[top_k - 1].z;
      nk = top_k;
      while (nk < V && a[nk].z == s->kth) ++nk;
    }
    if (use_p) {                       /* cumsum(softmax(sorted)) at T=1, shifted mask */
      const float c1 = (float)LOG2E;
      const float mc1 = m * c1;
      uint64_t S1 = 0;
      for (int64_t i = 0; i < nk; ++i) S1 += fix40(cexp2(fmaf(a[i].z, c1, -mc1)));
      uint64_t tpq = (uint64_t)((double)top_p * 4294967296.0);
      uint64_t thr = (uint64_t)(((u128)S1 * tpq) >> 32);
      uint64_t before = 0;
      int64_t last = 0;                /* rank 0 always kept */
      for (int64_t i = 0; i < nk; ++i) {
        if (i > 0 && before > thr) break;
        last = i;
        before += fix40(cexp2(fmaf(a[i].z, c1, -mc1)));
      }
      s->cut = a[last].z; s->jcut = a[last].j;
    }
    free(a);
  }
  uint64_t S = 0; int64_t nkept = 0;
  for (int64_t j = 0; j < V; ++j)
    if (row_kept(s, z[j], j)) { S += fix40(cexp2(fmaf(z[j], c, -s->mc))); ++nkept; }
  s->Sfix = S; s->n_kept = nkept;
  s->S32 = (float)S * 0x1p-40f;
  s->inv = 1.0f / s->S32;
}
static inline float row_e(const RowStats* s, const float* z, int64_t j, float c) {
  return row_kept(s, z[j], j) ? cexp2(fmaf(z[j], c, -s->mc)) : 0.0f;
}
static inline float row_prob(const RowStats* s, const float* z, int64_t j, float c) {
  return row_e(s, z, j, c) * s->inv;
}

/* sample() on a processed target row: greedy = utils/logits_processor.py:36,
 * otherwise inverse CDF on an injected uniform (SURVEY.md 8c restatement (i)). */
static int64_t sample_p_row(const RowStats* s, const float* z, int64_t V, float c, int greedy, float u) {
  if (greedy) {
    float best = -1.0f; int64_t bj = 0;
    for (int64_t j = 0; j < V; ++j) { float e = row_e(s, z, j, c); if (e > best) { best = e; bj = j; } }
    return bj;
  }
  uint64_t target = scale_u24(s->Sfix, u24_of(u));
  uint64_t cum = 0; int64_t lastpos = 0;
  for (int64_t j = 0; j < V; ++j) {
    uint64_t w = fix40(row_e(s, z, j, c));
    if (w) lastpos = j;
    cum += w;
    if (cum > target) return j;
  }
  return lastpos;
}
/* max_fn(p - q) then sample (sampling/speculative_decoding.py:10-19,168,171) */
static int64_t sample_residual(const RowStats* sp, const float* zp, const RowStats* sq, const float* zq,
                               int64_t V, float c, int greedy, float u, uint64_t rmin, int* fell_back) {
  uint64_t R = 0;
  float best = 0.0f; int64_t bj = -1;
  for (int64_t j = 0; j < V; ++j) {
    float r = row_prob(sp, zp, j, c) - row_prob(sq, zq, j, c);
    r = r > 0.0f ? r : 0.0f;
    R += fix60(r);
    if (r > best) { best = r; bj = j; }
  }
  *fell_back = 0;
  if (R <= rmin || bj < 0) { *fell_back = 1; return sample_p_row(sp, zp, V, c, greedy, u); }
  if (greedy) return bj;
  uint64_t target = scale_u24(R, u24_of(u));
  uint64_t cum = 0; int64_t lastpos = 0;
  for (int64_t j = 0; j < V; ++j) {
    float r = row_prob(sp, zp, j, c) - row_prob(sq, zq, j, c);
    r = r > 0.0f ? r : 0.0f;
    uint64_t w = fix60(r);
    if (w) lastpos = j;
    cum += w;
    if (cum > target) return j;
  }
  return lastpos;
}

/* One speculative verify step for B sequences (SURVEY.md 8a "exact per-row semantics").
 * Logits are fp32 (bf16/fp16 callers pass the exact .float() of their values).
 * Strides are in elements.  drf may be NULL only with F_NGRAM. */
ORACLE_API int oracle_verify(const float* tgt, const float* drf, int B, int gamma, int64_t V,
                             int64_t tsb, int64_t tsg, int64_t dsb, int64_t dsg,
                             const int64_t* draft_tokens, const float* u_accept, const float* u_sample,
                             float temperature, int top_k, float top_p, int greedy, int flags,
                             const int64_t* stop, int n_stop,
                             int32_t* n_acc, int64_t* next_tok, uint8_t* mask, float* p_tok, float* q_tok,
                             int32_t* first_stop) {
  const float c = (float)(LOG2E / (double)temperature);
  const uint64_t rmin = (flags & F_RESID_FALLBACK) ? 1152921ull : 0ull;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    RowStats* sp = (RowStats*)malloc(sizeof(RowStats) * (size_t)(gamma + 1));
    RowStats* sq = (RowStats*)malloc(sizeof(RowStats) * (size_t)(gamma + 1));
    int n = gamma;
    for (int i = 0; i < gamma; ++i) {
      const float* zp = tgt + b * tsb + i * tsg;
      row_stats(zp, V, c, top_k, top_p, &sp[i]);
      int64_t tok = draft_tokens[(int64_t)b * gamma + i];
      if (tok < 0 || tok >= V) { err = 1; tok = 0; }
      float p = row_prob(&sp[i], zp, tok, c);
      int acc;
      if (flags & F_NGRAM) {
        int64_t s = sample_p_row(&sp[i], zp, V, c, greedy, u_accept[(int64_t)b * gamma + i]);
        acc = (s == tok);
        p_tok[(int64_t)b * gamma + i] = p; q_tok[(int64_t)b * gamma + i] = 0.0f;
      } else {
        const float* zq = drf + b * dsb + i * dsg;
        row_stats(zq, V, c, top_k, top_p, &sq[i]);
        float q = row_prob(&sq[i], zq, tok, c);
        float u = u_accept[(int64_t)b * gamma + i];
        if (flags & F_ACCEPT_BATCHED) {
          double ap = (q <= 0.0f) ? 1.0 : fmin(1.0, (double)p / (double)q);
          acc = ((double)u < ap);
        } else {
          float frac = p / q;
          acc = !(u > frac);
        }
        p_tok[(int64_t)b * gamma + i] = p; q_tok[(int64_t)b * gamma + i] = q;
      }
      mask[(int64_t)b * gamma + i] = (uint8_t)acc;
      if (!acc && n == gamma) n = i;
    }
    n_acc[b] = n;
    /* sampling/speculative_decoding.py:150-152: torch.nonzero(eq(ids[1,n], stop[k,1]))[0,1] -- the first-LISTED stop
     * token that occurs among the accepted drafts decides (first position of it); the batched engine
     * (engine/infer_engine.py:310-312) breaks at the earliest accepted position holding any end token. */
    int fs = -1;
    if (flags & F_ACCEPT_BATCHED) {
      for (int i = 0; i < n && fs < 0; ++i)
        for (int k = 0; k < n_stop; ++k)
          if (draft_tokens[(int64_t)b * gamma + i] == stop[k]) { fs = i; break; }
    } else {
      for (int k = 0; k < n_stop && fs < 0; ++k)
        for (int i = 0; i < n; ++i)
          if (draft_tokens[(int64_t)b * gamma + i] == stop[k]) { fs = i; break; }
    }
    first_stop[b] = fs;
    float us = u_sample[b];
    int64_t x;
    if (n == gamma) {
      if (flags & F_NO_BONUS) x = -1;
      else {
        const float* zp = tgt + b * tsb + (int64_t)gamma * tsg;
        row_stats(zp, V, c, top_k, top_p, &sp[gamma]);
        x = sample_p_row(&sp[gamma], zp, V, c, greedy, us);
      }
    } else {
      const float* zp = tgt + b * tsb + n * tsg;
      if ((flags & F_NGRAM) || (flags & F_SKIP_ADJUST)) x = sample_p_row(&sp[n], zp, V, c, greedy, us);
      else {
        int fb;
        const float* zq = drf + b * dsb + n * dsg;
        x = sample_residual(&sp[n], zp, &sq[n], zq, V, c, greedy, us, rmin, &fb);
      }
    }
    next_tok[b] = x;
    free(sp); free(sq);
  }
  return err ? -1 : 0;
}

/* LogitsProcessor.__call__ materialised (utils/logits_processor.py:13-15) + row stats dump. */
ORACLE_API int oracle_process_probs(const float* z, int64_t rows, int64_t V, int64_t stride,
                                    float temperature, int top_k, float top_p, float* probs /*nullable*/,
                                    float* out_m, float* out_S32, uint64_t* out_Sfix, float* out_cut,
                                    int64_t* out_jcut, int64_t* out_nkept) {
  const float c = (float)(LOG2E / (double)temperature);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t r = 0; r < rows; ++r) {
    RowStats s;
    const float* zr = z + r * stride;
    row_stats(zr, V, c, top_k, top_p, &s);
    if (probs) for (int64_t j = 0; j < V; ++j) probs[r * V + j] = row_prob(&s, zr, j, c);
    if (out_m) out_m[r] = s.m;
    if (out_S32) out_S32[r] = s.S32;
    if (out_Sfix) out_Sfix[r] = s.Sfix;
    if (out_cut) out_cut[r] = s.kth > s.cut ? s.kth : s.cut;
    if (out_jcut) out_jcut[r] = s.jcut;
    if (out_nkept) out_nkept[r] = s.n_kept;
  }
  return 0;
}

/* fused processor+sample for one row per call site (AR / drafter step):
 * token ~ processor(logits) via greedy or inverse CDF; also returns prob of the token. */
ORACLE_API int oracle_sample_rows(const float* z, int64_t rows, int64_t V, int64_t stride, float temperature,
                                  int top_k, float top_p, int greedy, const float* u, int64_t* tok, float* ptok) {
  const float c = (float)(LOG2E / (double)temperature);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t r = 0; r < rows; ++r) {
    RowStats s;
    const float* zr = z + r * stride;
    row_stats(zr, V, c, top_k, top_p, &s);
    int64_t x = sample_p_row(&s, zr, V, c, greedy, u ? u[r] : 0.0f);
    tok[r] = x;
    if (ptok) ptok[r] = row_prob(&s, zr, x, c);
  }
  return 0;
}

/* sample() on already-materialised probabilities (utils/logits_processor.py:35-36,48-49 shape
 * contract): greedy argmax first index, else inverse CDF with weights (uint64)(p*2^40). */
ORACLE_API int oracle_sample_probs(const float* p, int64_t rows, int64_t V, int greedy, const float* u, int64_t* tok) {
  for (int64_t r = 0; r < rows; ++r) {
    const float* pr = p + r * V;
    if (greedy) {
      float best = -INFINITY; int64_t bj = 0;
      for (int64_t j = 0; j < V; ++j) if (pr[j] > best) { best = pr[j]; bj = j; }
      tok[r] = bj; continue;
    }
    uint64_t S = 0;
    for (int64_t j = 0; j < V; ++j) S += fix40(pr[j] > 0.0f ? pr[j] : 0.0f);
    uint64_t target = scale_u24(S, u24_of(u[r]));
    uint64_t cum = 0; int64_t lastpos = 0, x = -1;
    for (int64_t j = 0; j < V; ++j) {
      uint64_t w = fix40(pr[j] > 0.0f ? pr[j] : 0.0f);
      if (w) lastpos = j;
      cum += w;
      if (cum > target) { x = j; break; }
    }
    tok[r] = x < 0 ? lastpos : x;
  }
  return 0;
}

/* ---- Philox4x32-10 keyed uniforms (production RNG contract) ---- */
static inline void philox_round(uint32_t* c, uint32_t k0, uint32_t k1) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
static inline uint32_t philox_word0(uint64_t seed, uint64_t offset, uint32_t seq, uint32_t lane) {
  uint32_t c[4] = {(uint32_t)offset, (uint32_t)(offset >> 32), seq, lane};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c[0];
}
/* u_accept[b,i] = stream 0, u_sample[b] = stream 1 ; seq ids are GLOBAL (seq0 + b) */
ORACLE_API int oracle_philox_uniform(uint64_t seed, uint64_t offset, int64_t seq0, int B, int gamma,
                                     float* u_accept, float* u_sample) {
  for (int b = 0; b < B; ++b) {
    uint32_t s = (uint32_t)(seq0 + b);
    for (int i = 0; i < gamma; ++i)
      u_accept[(int64_t)b * gamma + i] = (float)(philox_word0(seed, offset, s, (uint32_t)i) >> 8) * 0x1p-24f;
    u_sample[b] = (float)(philox_word0(seed, offset, s, 0x10000u) >> 8) * 0x1p-24f;
  }
  return 0;
}

/* test helpers: expose the canonical primitives */
ORACLE_API float oracle_cexp2(float t) { return cexp2(t); }
ORACLE_API void oracle_cexp2_array(const float* t, float* out, int64_t n) { for (int64_t i = 0; i < n; ++i) out[i] = cexp2(t[i]); }
ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORACLE_API void oracle_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- KV rollback (utils/caching.py:27-55 applied per sequence) ----
 * Static cache layout [B, H, S_max, D] per layer tensor; sequence b drops its last
 * discard[b] valid positions; the dropped region is zero-filled so that the valid
 * prefix equals prune_tuple_cache's view tensor[:, :, :-n, :] of that sequence. */
ORACLE_API int oracle_prune_kv(uint8_t** tensors, int n_tensors, int B, int H, int64_t S_max, int64_t D,
                               int elem_bytes, int32_t* seq_lens, const int32_t* discard, int zero_fill) {
  for (int b = 0; b < B; ++b) {
    int32_t old = seq_lens[b];
    int32_t d = discard[b];
    if (d < 0) d = 0;
    if (d > old) d = old;
    int32_t nw = old - d;
    if (zero_fill)
      for (int t = 0; t < n_tensors; ++t)
        for (int h = 0; h < H; ++h) {
          uint8_t* base = tensors[t] + (((int64_t)b * H + h) * S_max) * D * elem_bytes;
          memset(base + (int64_t)nw * D * elem_bytes, 0, (size_t)((int64_t)(old - nw) * D * elem_bytes));
        }
    seq_lens[b] = nw;
  }
  return 0;
}
