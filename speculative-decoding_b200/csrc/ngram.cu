// ngram.cu -- device n-gram tables for n-gram-assisted speculative decoding (sm_100a).
// Replaces NGramStorage / OneLevelNGramStorage (ngram_assisted/ngram_storage.py:73-249), which keep
// Python dict-of-dict counts: here every table is an open-addressing hash table in HBM holding, per
// context (gram), the arg-max-count token (strict '>' keeps the incumbent, :127,:220) and a second
// table of (gram, token) -> count.  One logical table per table id (= per sequence, the layout the
// batched configs need); one thread owns one table and applies its sequences' updates in batch
// order, so results equal the reference's sequential dict updates exactly (integer work, no atomics).
// The gamma chained next_token() probes of ngram_assisted/ngram_assisted.py:95-99 run in one launch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/specdec_b200.h"
#include "canon.cuh"  // philox_word0 (fallback tokens of unknown contexts)

#define NG_MAXCTX 8

struct GramSlot {  // 48 bytes
  int len;         // 0 = empty, else context length j
  int tok[NG_MAXCTX];
  int best_tok;
  int best_cnt;
  int pad;
};
struct CountSlot {  // 16 bytes
  int gram;         // gram slot index + 1, 0 = empty
  int tok;
  int cnt;
  int pad;
};
struct specdec_ngram {
  int n, vocab, n_tables, G, C, one_level;
  GramSlot* grams;
  CountSlot* counts;
  int* status;  // [0] overflow flag, [1] max grams used
  int* used;    // per table gram count
  unsigned long long seed, calls;  // fallback tokens of unknown contexts: Philox(seed; call number, sequence, position)
};

namespace specdec {

__device__ __forceinline__ unsigned ng_hash(const long long* ctx, int j) {
  unsigned h = 2166136261u ^ (unsigned)j;
  for (int i = 0; i < j; ++i) { h ^= (unsigned)ctx[i]; h *= 16777619u; h ^= h >> 13; }
  return h * 2654435761u;
}
// returns slot index (within table) or -1; insert!=0 creates the gram (best_tok unset = -1)
__device__ int ng_find_gram(GramSlot* g, int G, const long long* ctx, int j, int insert, int* created) {
  unsigned h = ng_hash(ctx, j) % (unsigned)G;
  for (int probe = 0; probe < G; ++probe) {
    GramSlot& s = g[h];
    if (s.len == 0) {
      if (!insert) return -1;
      s.len = j;
      for (int i = 0; i < j; ++i) s.tok[i] = (int)ctx[i];
      s.best_tok = -1; s.best_cnt = 0;
      *created = 1;
      return (int)h;
    }
    if (s.len == j) {
      bool eq = true;
      for (int i = 0; i < j; ++i) eq = eq && (s.tok[i] == (int)ctx[i]);
      if (eq) return (int)h;
    }
    h = (h + 1 == (unsigned)G) ? 0u : h + 1;
  }
  return -2;  // table full
}
// increments count of (gram, tok); returns new count or -2 if full
__device__ int ng_bump(CountSlot* c, int C, int gram, int tok) {
  unsigned h = ((unsigned)gram * 2654435761u ^ (unsigned)tok * 40503u) % (unsigned)C;
  for (int probe = 0; probe < C; ++probe) {
    CountSlot& s = c[h];
    if (s.gram == 0) { s.gram = gram + 1; s.tok = tok; s.cnt = 1; return 1; }
    if (s.gram == gram + 1 && s.tok == tok) return ++s.cnt;
    h = (h + 1 == (unsigned)C) ? 0u : h + 1;
  }
  return -2;
}
// counts[j][gram][token] += 1 with the reference's arg-max rule (ngram_storage.py:209-221 / :235-245)
__device__ void ng_observe(GramSlot* g, int G, CountSlot* c, int C, const long long* ctx, int j,
                           const long long* toks, int m, int* status, int* used) {
  int created = 0;
  const int gi = ng_find_gram(g, G, ctx, j, 1, &created);
  if (gi < 0) { status[0] = 1; return; }
  if (created) { g[gi].best_tok = (int)toks[0]; ++(*used); }
  for (int t = 0; t < m; ++t) {
    const int tok = (int)toks[t];
    if (tok < 0) continue;  // padding of a batched update whose rows carry different numbers of tokens
    const int cnt = ng_bump(c, C, gi, tok);
    if (cnt < 0) { status[0] = 1; return; }
    if (tok == g[gi].best_tok) g[gi].best_cnt = cnt;              // incumbent's own count moves
    else if (cnt > 1 && cnt > g[gi].best_cnt) { g[gi].best_tok = tok; g[gi].best_cnt = cnt; }
    // cnt == 1 (first sighting): the reference does not compare (ngram_storage.py:215-216)
  }
}

// mode 0 = initialize (slide over the whole prefix), 1 = update (last contexts only)
__global__ void ngram_write_kernel(specdec_ngram t, const long long* ids, const int* lens, const int* table_ids,
                                   int B, long long max_len, const long long* next_tokens, int m, int mode) {
  const int tab = blockIdx.x * blockDim.x + threadIdx.x;
  if (tab >= t.n_tables) return;
  GramSlot* g = t.grams + (size_t)tab * t.G;
  CountSlot* c = t.counts + (size_t)tab * t.C;
  int* used = t.used + tab;
  for (int b = 0; b < B; ++b) {
    if ((table_ids ? table_ids[b] : 0) != tab) continue;
    const long long* seq = ids + (size_t)b * max_len;
    const int len = lens[b];
    if (mode == 1) {
      const long long* nt = next_tokens + (size_t)b * m;
      if (t.one_level) {  // OneLevelNGramStorage.update: needs len >= n (ngram_storage.py:110)
        if (len < t.n) continue;
        ng_observe(g, t.G, c, t.C, seq + len - (t.n - 1), t.n - 1, nt, m, t.status, used);
      } else {            // NGramStorage.update: j = min(n-1,len) .. 2 (ngram_storage.py:200)
        if (len < 1) continue;
        for (int j = min(t.n - 1, len); j > 1; --j)
          ng_observe(g, t.G, c, t.C, seq + len - j, j, nt, m, t.status, used);
      }
    } else {
      if (t.one_level) {  // ngram_storage.py:132-146
        for (int i = 0; i + t.n - 1 < len; ++i)
          ng_observe(g, t.G, c, t.C, seq + i, t.n - 1, seq + i + t.n - 1, 1, t.status, used);
      } else {            // ngram_storage.py:225-245
        for (int i = 0; i < len; ++i)
          for (int j = min(t.n - 1, i); j > 1; --j)
            ng_observe(g, t.G, c, t.C, seq + i - j, j, seq + i, 1, t.status, used);
      }
    }
  }
  atomicMax(&t.status[1], *used);
}

// read-only probe of the (gram, token) count table: the count, 0 if never seen
__device__ int ng_count(const CountSlot* c, int C, int gram, int tok) {
  unsigned h = ((unsigned)gram * 2654435761u ^ (unsigned)tok * 40503u) % (unsigned)C;
  for (int probe = 0; probe < C; ++probe) {
    const CountSlot& s = c[h];
    if (s.gram == 0) return 0;
    if (s.gram == gram + 1 && s.tok == tok) return s.cnt;
    h = (h + 1 == (unsigned)C) ? 0u : h + 1;
  }
  return 0;
}

// gamma chained lookups per sequence (ngram_assisted.py:95-99 -> ngram_storage.py:164-179 / :83-96).
// One WARP per sequence: lane j probes context length j, so the back-off levels of one next_token() call are looked up
// side by side (their hash probes are dependent global loads: the chain of gamma calls is latency bound) and the
// longest hit wins -- the order the reference's `for j in range(min(n-1, len), 1, -1)` loop tries them in.
__global__ void __launch_bounds__(128) ngram_lookup_kernel(specdec_ngram t, const long long* ids, const int* lens, const int* table_ids,
                                                           int B, long long max_len, int gamma, const long long* fallback,
                                                           long long* drafts, unsigned char* known) {
  __shared__ long long s_win[4][NG_MAXCTX + 64];  // last n-1 real tokens followed by the drafts made so far
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int b = blockIdx.x * 4 + wib;
  if (b >= B) return;
  const int tab = table_ids ? table_ids[b] : 0;
  const GramSlot* g = t.grams + (size_t)tab * t.G;
  const long long* seq = ids + (size_t)b * max_len;
  const int len0 = lens[b];
  long long* win = s_win[wib];
  const int keep = min(t.n - 1, len0);
  if (lane < keep) win[lane] = seq[len0 - keep + lane];
  __syncwarp();
  int wl = keep;
  for (int k = 0; k < gamma; ++k) {
    const int len = len0 + k;
    int dummy = 0, gi = -1;
    const int jmax = min(t.n - 1, len);
    const bool mine = t.one_level ? (lane == t.n - 1 && len >= t.n - 1) : (lane >= 2 && lane <= jmax);
    if (mine) gi = ng_find_gram((GramSlot*)g, t.G, win + wl - lane, lane, 0, &dummy);
    const unsigned hits = __ballot_sync(0xffffffffu, gi >= 0);
    long long out;
    unsigned char kn = 0;
    if (hits) {
      const int src = 31 - __clz(hits);  // longest context that is known
      const int bt = (gi >= 0) ? g[gi].best_tok : 0;
      out = (long long)__shfl_sync(0xffffffffu, bt, src);
      kn = 1;
    } else if (fallback) {
      out = fallback[(size_t)b * gamma + k];
    } else {  // the reference draws torch.randint(vocab_size) (ngram_storage.py:84,165)
      out = (long long)(philox_word0(t.seed, t.calls, (unsigned)b, (unsigned)k) % (unsigned)t.vocab);
    }
    if (lane == 0) {
      drafts[(size_t)b * gamma + k] = out;
      known[(size_t)b * gamma + k] = kn;
    }
    // slide the window
    __syncwarp();
    if (wl == NG_MAXCTX + 63) {
      long long v = (lane + 1 < wl) ? win[lane + 1] : 0, v2 = (lane + 33 < wl) ? win[lane + 33] : 0, v3 = (lane + 65 < wl) ? win[lane + 65] : 0;
      __syncwarp();
      if (lane + 1 < wl) win[lane] = v;
      if (lane + 33 < wl) win[lane + 32] = v2;
      if (lane + 65 < wl) win[lane + 64] = v3;
      --wl;
      __syncwarp();
    }
    if (lane == 0) win[wl] = out;
    ++wl;
    __syncwarp();
  }
}

// has_gram (ngram_storage.py:98-106 / :181-193), exact: the reference takes the LAST j tokens of `ngram` (its final
// token included) as the context and asks whether that same final token has ever been counted after it.
__global__ void ngram_has_kernel(specdec_ngram t, const long long* ids, const int* lens, const int* table_ids, int B,
                                 long long max_len, unsigned char* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int tab = table_ids ? table_ids[b] : 0;
  const GramSlot* g = t.grams + (size_t)tab * t.G;
  const CountSlot* c = t.counts + (size_t)tab * t.C;
  const long long* seq = ids + (size_t)b * max_len;
  const int len = lens[b];
  unsigned char res = 0;
  int dummy = 0;
  if (t.one_level) {
    if (len >= t.n) {
      const int gi = ng_find_gram((GramSlot*)g, t.G, seq + len - (t.n - 1), t.n - 1, 0, &dummy);
      if (gi >= 0 && ng_count(c, t.C, gi, (int)seq[len - 1]) > 0) res = 1;
    }
  } else if (len >= 1) {
    for (int j = min(t.n - 1, len); j > 1 && !res; --j) {
      const int gi = ng_find_gram((GramSlot*)g, t.G, seq + len - j, j, 0, &dummy);
      if (gi >= 0 && ng_count(c, t.C, gi, (int)seq[len - 1]) > 0) res = 1;
    }
  }
  out[b] = res;
}

}  // namespace specdec

extern "C" {

int specdec_ngram_create(specdec_ngram_t** out, int n, int vocab_size, int n_tables, int grams_per_table,
                         int counts_per_table, int one_level) {
  if (!out || n < 2 || n - 1 > NG_MAXCTX || n_tables <= 0 || grams_per_table <= 0 || counts_per_table <= 0)
    return SPECDEC_ERR_ARG;
  specdec_ngram* t = (specdec_ngram*)calloc(1, sizeof(specdec_ngram));
  if (!t) return SPECDEC_ERR_ARG;
  t->n = n; t->vocab = vocab_size; t->n_tables = n_tables; t->G = grams_per_table; t->C = counts_per_table;
  t->one_level = one_level ? 1 : 0;
  cudaError_t e;
  if ((e = cudaMalloc(&t->grams, sizeof(GramSlot) * (size_t)n_tables * t->G)) != cudaSuccess) { free(t); return (int)e; }
  if ((e = cudaMalloc(&t->counts, sizeof(CountSlot) * (size_t)n_tables * t->C)) != cudaSuccess) { cudaFree(t->grams); free(t); return (int)e; }
  if ((e = cudaMalloc(&t->status, sizeof(int) * 2)) != cudaSuccess) { cudaFree(t->grams); cudaFree(t->counts); free(t); return (int)e; }
  if ((e = cudaMalloc(&t->used, sizeof(int) * (size_t)n_tables)) != cudaSuccess) { cudaFree(t->grams); cudaFree(t->counts); cudaFree(t->status); free(t); return (int)e; }
  *out = t;
  // the tables are zero before create returns, whatever stream the caller works on afterwards
  const int rc = specdec_ngram_reset(t, nullptr);
  if (rc) return rc;
  return (int)cudaStreamSynchronize(nullptr);
}
int specdec_ngram_seed(specdec_ngram_t* t, uint64_t seed) {
  if (!t) return SPECDEC_ERR_ARG;
  t->seed = seed; t->calls = 0;
  return 0;
}
int specdec_ngram_has_gram(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids, int B,
                           int64_t max_len, uint8_t* out, specdec_stream_t stream) {
  if (!t || !ids || !lens || !out || B < 0) return SPECDEC_ERR_ARG;
  if (B == 0) return 0;
  specdec::ngram_has_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*t, (const long long*)ids, lens, table_ids, B, max_len, out);
  return (int)cudaGetLastError();
}
int specdec_ngram_destroy(specdec_ngram_t* t) {
  if (!t) return 0;
  cudaFree(t->grams); cudaFree(t->counts); cudaFree(t->status); cudaFree(t->used);
  free(t);
  return 0;
}
int specdec_ngram_reset(specdec_ngram_t* t, specdec_stream_t stream) {
  if (!t) return SPECDEC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if ((e = cudaMemsetAsync(t->grams, 0, sizeof(GramSlot) * (size_t)t->n_tables * t->G, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(t->counts, 0, sizeof(CountSlot) * (size_t)t->n_tables * t->C, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(t->status, 0, sizeof(int) * 2, st)) != cudaSuccess) return (int)e;
  return (int)cudaMemsetAsync(t->used, 0, sizeof(int) * (size_t)t->n_tables, st);
}
int specdec_ngram_initialize(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                             int B, int64_t max_len, specdec_stream_t stream) {
  if (!t || !ids || !lens || B < 0) return SPECDEC_ERR_ARG;
  if (B == 0) return 0;
  specdec::ngram_write_kernel<<<(t->n_tables + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      *t, (const long long*)ids, lens, table_ids, B, max_len, nullptr, 0, 0);
  return (int)cudaGetLastError();
}
int specdec_ngram_update(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                         int B, int64_t max_len, const int64_t* next_tokens, int m, specdec_stream_t stream) {
  if (!t || !ids || !lens || !next_tokens || B < 0 || m <= 0) return SPECDEC_ERR_ARG;
  if (B == 0) return 0;
  specdec::ngram_write_kernel<<<(t->n_tables + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
      *t, (const long long*)ids, lens, table_ids, B, max_len, (const long long*)next_tokens, m, 1);
  return (int)cudaGetLastError();
}
int specdec_ngram_lookup_chain(specdec_ngram_t* t, const int64_t* ids, const int32_t* lens, const int32_t* table_ids,
                               int B, int64_t max_len, int gamma, const int64_t* fallback, int64_t* drafts,
                               uint8_t* known, specdec_stream_t stream) {
  if (!t || !ids || !lens || !drafts || !known || B < 0 || gamma < 0 || gamma > 64) return SPECDEC_ERR_ARG;
  if (B == 0 || gamma == 0) return 0;
  specdec::ngram_lookup_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      *t, (const long long*)ids, lens, table_ids, B, max_len, gamma, (const long long*)fallback,
      (long long*)drafts, known);
  if (!fallback) ++t->calls;  // every call without caller-supplied fallbacks draws fresh ones
  return (int)cudaGetLastError();
}
int specdec_ngram_status(specdec_ngram_t* t, int32_t* host_out2) {
  if (!t || !host_out2) return SPECDEC_ERR_ARG;
  return (int)cudaMemcpy(host_out2, t->status, sizeof(int) * 2, cudaMemcpyDeviceToHost);
}

}  // extern "C"
