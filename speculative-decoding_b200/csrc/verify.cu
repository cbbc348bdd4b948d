// verify.cu -- fused speculative-sampling verify for sm_100a (B200).
//
// Two kernels per verify step:
//   rowstats_kernel : one CTA per (sequence, position, model) logit row.  Replaces
//                     LogitsProcessor.__call__ (utils/logits_processor.py:13-15) incl. top-k (:59-63)
//                     and nucleus (:73-81, :92-103) WITHOUT materialising probabilities: per row it
//                     emits {max, 1/sum, cut value, cut index} (32 bytes).
//   decide_kernel   : one CTA per sequence.  Accept test (sampling/speculative_decoding.py:139-145 or
//                     engine/infer_engine.py:303-305), stop scan (:150-155), then ONE more sweep over
//                     the single row (pair) that needs it: residual max(0,p-q) renormalisation +
//                     inverse-CDF resample (:10-19,168,171) or the bonus row (:158-160).
// All arithmetic is the canonical fp32/integer arithmetic of canon.cuh, so results are bit-exact
// against the CPU oracle for every launch geometry.
#include <string.h>
#include <mutex>
#include "canon.cuh"
#include "../../include/specdec_b200.h"

namespace specdec {

constexpr int NT = 1024;       // threads per CTA (32 warps)
constexpr int RS_NT = 512;     // threads per CTA of rowstats_kernel (2 CTAs / SM)
constexpr int CAP = 4096;      // top-k / nucleus candidate capacity per row (shared memory)
constexpr int WARP_SELECT_MAX = 1024;  // candidate sets up to this size are resolved by one warp
constexpr int MAXPART = 4096;  // warp-vector partial sums per row: V <= MAXPART*256
constexpr size_t CAND_SMEM = (size_t)CAP * (sizeof(float) + sizeof(int) + sizeof(u64));

struct RowOut {  // 32 bytes per logit row
  float m, mc, inv, cut;
  int jcut, flags;
  u64 Sfix;
};

struct RowJob {
  const void* tgt;
  const void* drf;
  long long tsb, tsg, dsb, dsg;  // element strides
  int nT, nD;                    // target / draft rows per sequence
  int V;
  float c, c1;
  int top_k;  // 0 = off
  int use_p;
  u64 tpq;  // floor(top_p * 2^32)
  long long R;
  RowOut* out;
  int skip_resolved;  // rowstats_kernel: skip rows whose RowOut is already flagged exact
  int pre_stats;      // nucleus_fast_kernel: max / MUFU mass / candidate threshold come from rowfast_tma_kernel<DT, 1>
  int2* klist;        // nullable: [R][KL_MAX] (logit bits, index) of the kept tokens of rows whose kept set is small;
                      // the count sits in RowOut.flags bits 8..15 (0 = no list).  Feeds sample_lists_kernel.
  int* n_unres;       // nullable: rows rowsel_tma_kernel left unresolved (zeroed at the head of the call); the exact
                      // fallback kernels return at once when it is 0 instead of polling every row's flag
};
constexpr int KL_MAX = 64;

template <int DT>
__device__ __forceinline__ const void* row_ptr(const RowJob& job, long long r) {
  const int rps = job.nT + job.nD;
  const long long b = r / rps;
  const int k = (int)(r - b * rps);
  const size_t es = (DT == DT_F32) ? 4 : 2;
  if (k < job.nT) return (const char*)job.tgt + (size_t)(b * job.tsb + (long long)k * job.tsg) * es;
  return (const char*)job.drf + (size_t)(b * job.dsb + (long long)(k - job.nT) * job.dsg) * es;
}

// Visit every 8-element vector of a row: f(x[8], j0).  4 independent vector loads in flight.
template <int DT, int NTH, typename F>
__device__ __forceinline__ void sweep_range(const void* row, int V, bool aligned, int v_begin, int v_end, F f) {
  int v = v_begin + threadIdx.x;
  for (; v + 3 * NTH < v_end; v += 4 * NTH) {
    float x0[8], x1[8], x2[8], x3[8];
    load8<DT>(row, v, V, aligned, x0);
    load8<DT>(row, v + NTH, V, aligned, x1);
    load8<DT>(row, v + 2 * NTH, V, aligned, x2);
    load8<DT>(row, v + 3 * NTH, V, aligned, x3);
    f(x0, v * 8);
    f(x1, (v + NTH) * 8);
    f(x2, (v + 2 * NTH) * 8);
    f(x3, (v + 3 * NTH) * 8);
  }
  for (; v < v_end; v += NTH) {
    float x[8];
    load8<DT>(row, v, V, aligned, x);
    f(x, v * 8);
  }
}
template <int DT, typename F>
__device__ __forceinline__ void sweep(const void* row, int V, bool aligned, F f) {
  const int NV = (V + 7) >> 3;
  const int NTB = cta_nthreads();
  int v = threadIdx.x;
  for (; v + 3 * NTB < NV; v += 4 * NTB) {
    float x0[8], x1[8], x2[8], x3[8];
    load8<DT>(row, v, V, aligned, x0);
    load8<DT>(row, v + NTB, V, aligned, x1);
    load8<DT>(row, v + 2 * NTB, V, aligned, x2);
    load8<DT>(row, v + 3 * NTB, V, aligned, x3);
    f(x0, v * 8);
    f(x1, (v + NTB) * 8);
    f(x2, (v + 2 * NTB) * 8);
    f(x3, (v + 3 * NTB) * 8);
  }
  for (; v < NV; v += NTB) {
    float x[8];
    load8<DT>(row, v, V, aligned, x);
    f(x, v * 8);
  }
}

// ---------------------------------------------------------------------------------------------
// top-k / nucleus cut selection, generic over the element source (candidate list or full row)
// ---------------------------------------------------------------------------------------------
struct CandSrc {
  float* cz;
  int* cj;
  u64* cw;
  int n;
  float c1, mc1;
  __device__ void prepare(unsigned kthkey) const {  // T=1 masses of the top-k-kept candidates
    for (int i = threadIdx.x; i < n; i += blockDim.x) cw[i] = (fkey(cz[i]) >= kthkey) ? fix40(cweight(cz[i], c1, mc1)) : 0ull;
    __syncthreads();
  }
  template <typename F>
  __device__ void each(unsigned, bool, F f) const {
    for (int i = threadIdx.x; i < n; i += blockDim.x) f(cz[i], fkey(cz[i]), cj[i], cw[i]);
  }
};
template <int DT>
struct RowSrc {
  const void* row;
  int V;
  bool aligned;
  float c1, mc1;
  __device__ void prepare(unsigned) const {}
  template <typename F>
  __device__ void each(unsigned kthkey, bool need_w, F f) const {
    sweep<DT>(row, V, aligned, [&](const float(&x)[8], int j0) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (j0 + k < V) {
          const unsigned key = fkey(x[k]);
          const u64 w = (need_w && key >= kthkey) ? fix40(cweight(x[k], c1, mc1)) : 0ull;
          f(x[k], key, j0 + k, w);
        }
    });
  }
};

// Finds (cut, jcut, Sfix) for the masked modes.  S1_full: T=1 mass of the whole row (pure nucleus
// only; with top-k it is recomputed over the kept set).  Block-uniform control flow.
template <typename Src>
__device__ void select_cut(const Src& src, int V, int top_k, int use_p, u64 tpq, u64 S1_full, float c, float mc,
                           u64* sh64, unsigned* shu, float& cut_out, int& jcut_out, u64& Sfix_out) {
  unsigned kthkey = 0;
  if (top_k > 0) {  // largest key K with count(key >= K) >= top_k   (utils/logits_processor.py:61)
    unsigned K = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned tr = K | (1u << bit);
      u64 cnt = 0;
      src.each(0u, false, [&](float, unsigned key, int, u64) { cnt += (key >= tr) ? 1u : 0u; });
      cnt = block_sum_u64(cnt, sh64);
      if (cnt >= (u64)top_k) K = tr;
    }
    kthkey = K;
  }
  src.prepare(kthkey);
  unsigned cutkey = kthkey;
  int jcut = V;
  if (use_p) {
    u64 S1 = S1_full;
    if (top_k > 0) {
      u64 loc = 0;
      src.each(kthkey, true, [&](float, unsigned, int, u64 w) { loc += w; });
      S1 = block_sum_u64(loc, sh64);
    }
    const u64 thr = scale_q32(S1, tpq);
    // v' = max key whose strictly-greater mass still exceeds thr
    unsigned vp = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned tr = vp | (1u << bit);
      u64 loc = 0;
      src.each(kthkey, true, [&](float, unsigned key, int, u64 w) { loc += (key > tr) ? w : 0ull; });
      const u64 G = block_sum_u64(loc, sh64);
      if (G > thr) vp = tr;
    }
    unsigned lk = 0xFFFFFFFFu;
    u64 locG = 0;
    src.each(kthkey, false, [&](float, unsigned key, int, u64) {
      if (key > vp && key >= kthkey) lk = min(lk, key);
    });
    cutkey = block_min_u32(lk, shu);
    u64 cntc = 0;
    src.each(kthkey, true, [&](float, unsigned key, int, u64 w) {
      locG += (key > cutkey) ? w : 0ull;
      cntc += (key == cutkey) ? 1u : 0u;
    });
    const u64 Gc = block_sum_u64(locG, sh64);
    cntc = block_sum_u64(cntc, sh64);
    const u64 wc = fix40(cweight(fkey_inv(cutkey), src.c1, src.mc1));
    u64 mkeep = cntc;
    if (wc > 0 && thr >= Gc) {
      const u64 q = (thr - Gc) / wc + 1ull;
      mkeep = q < cntc ? q : cntc;
    }
    if (mkeep < cntc) {  // keep the first mkeep ties in ascending index order
      int lo = 0, hi = V - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        u64 cnt = 0;
        src.each(kthkey, false, [&](float, unsigned key, int j, u64) { cnt += (key == cutkey && j <= mid) ? 1u : 0u; });
        cnt = block_sum_u64(cnt, sh64);
        if (cnt >= mkeep) hi = mid; else lo = mid + 1;
      }
      jcut = lo;
    }
  }
  u64 loc = 0;
  src.each(0u, false, [&](float z, unsigned key, int j, u64) {
    if (key > cutkey || (key == cutkey && j <= jcut)) loc += fix40(cweight(z, c, mc));
  });
  Sfix_out = block_sum_u64(loc, sh64);
  cut_out = fkey_inv(cutkey);
  jcut_out = jcut;
}

// ---------------------------------------------------------------------------------------------
// single-warp cut selection for small candidate sets (executed by warp 0 while the other warps
// wait at a barrier and the SM's second CTA keeps streaming).  Same result as select_cut.
// Both selections are "the largest key K such that sum_{key_i >= K} w_i > t0" (w = 1, t0 = k-1 for
// top-k; w = T=1 mass, t0 = thr for the nucleus cut), found bit by bit over the bits in which the
// candidates actually differ (bf16-origin logits: <= ~12 of 32).
// ---------------------------------------------------------------------------------------------
// Warp-cooperative compaction of the elements of one 8-element vector per lane that lie in [th, up):
// per-lane hit mask -> warp prefix sum -> ONE shared-memory atomic per warp -> predicated stores.
// (x holds -inf for elements past the row end.)  Must be called by all 32 lanes.
__device__ __forceinline__ void compact_vec(const float (&x)[8], int v, float th, float up, int V, float* cz, int* cj,
                                            int* s_count, int cap) {
  const int lane = threadIdx.x & 31;
  // hits are usually rare: one vote on the vector maximum skips the per-element work
  const float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
  if (!__any_sync(0xffffffffu, vm >= th)) return;
  unsigned mask = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if ((x[k] >= th) && (x[k] < up) && (v * 8 + k < V)) mask |= 1u << k;
  if (!__any_sync(0xffffffffu, mask != 0)) return;
  const int cnt = __popc(mask);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  int base = 0;
  if (lane == 31) base = atomicAdd(s_count, incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  int pos = base + incl - cnt;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if ((mask >> k) & 1u) {
      if (pos < cap) { cz[pos] = x[k]; cj[pos] = v * 8 + k; }
      ++pos;
    }
}

// reduction group: warp 0 (BLK = false) or the whole CTA (BLK = true)
template <bool BLK>
struct Grp {
  u64* sh64;
  unsigned* shu;
  __device__ __forceinline__ int first() const { return BLK ? threadIdx.x : (threadIdx.x & 31); }
  __device__ __forceinline__ int stride() const { return BLK ? blockDim.x : 32; }
  __device__ __forceinline__ u64 sum(u64 v) const { return BLK ? block_sum_u64(v, sh64) : warp_sum_u64(v); }
  __device__ __forceinline__ unsigned orr(unsigned v) const {
    v = __reduce_or_sync(0xffffffffu, v);
    if (!BLK) return v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) shu[w] = v;
    __syncthreads();
    unsigned t = (lane < nw) ? shu[lane] : 0u;
    t = __reduce_or_sync(0xffffffffu, t);
    __syncthreads();
    return t;
  }
  __device__ __forceinline__ void sync() const { if (BLK) __syncthreads(); else __syncwarp(); }
};

// largest key K with sum_{key_i >= K} w_i > t0, searched over the bits in which the keys differ
template <bool MASS, bool BLK>
__device__ __forceinline__ unsigned group_select_key(const Grp<BLK>& gp, const float* cz, const u64* cw, int m,
                                                     unsigned common, unsigned vary, u64 t0) {
  unsigned K = common;
  for (int bit = 31; bit >= 0; --bit) {
    if (!((vary >> bit) & 1u)) continue;
    const unsigned tr = K | (1u << bit);
    u64 loc = 0;
    for (int i = gp.first(); i < m; i += gp.stride()) {
      if (fkey(cz[i]) >= tr) loc += MASS ? cw[i] : 1ull;
    }
    loc = gp.sum(loc);
    if (loc > t0) K = tr;
  }
  return K;
}

// Cut selection among n candidates in shared memory.  Band mode (pure nucleus): the candidates are the
// elements of one value band; G_off = exact T=1 mass of everything above the band (all kept) and
// S_above = their tempered weight sum.
template <bool BLK>
__device__ void select_cut_group(const Grp<BLK>& gp, float* cz, int* cj, u64* cw, int n, int V, int top_k, int use_p,
                                 u64 tpq, u64 S1_full, u64 G_off, u64 S_above, float c, float mc, float c1, float mc1,
                                 float& cut_out, int& jcut_out, u64& Sfix_out, u64 thr_lo = 0, u64 thr_hi = 0,
                                 int* ambiguous = nullptr, int* m_out = nullptr) {
  const int lane = threadIdx.x & 31;
  const unsigned k0 = fkey(cz[0]);
  unsigned vary = 0;
  for (int i = gp.first(); i < n; i += gp.stride()) vary |= fkey(cz[i]) ^ k0;
  vary = gp.orr(vary);
  unsigned kthkey = 0;
  int m = n;
  if (top_k > 0) {  // (warp group only: top-k candidate sets are small)
    kthkey = group_select_key<false, BLK>(gp, cz, cw, n, k0 & ~vary, vary, (u64)top_k - 1ull);
    m = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const float z = (i < n) ? cz[i] : 0.0f;
      const int j = (i < n) ? cj[i] : 0;
      const bool keep = (i < n) && (fkey(z) >= kthkey);
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) { const int pos = m + __popc(bal & ((1u << lane) - 1u)); cz[pos] = z; cj[pos] = j; }
      m += __popc(bal);
      __syncwarp();
    }
  }
  unsigned cutkey = kthkey;
  int jcut = V;
  if (use_p) {
    u64 loc = 0;
    unsigned v2 = 0;
    const unsigned k1 = fkey(cz[0]);
    for (int i = gp.first(); i < m; i += gp.stride()) {
      const u64 w = fix40(cweight(cz[i], c1, mc1));
      cw[i] = w;
      loc += w;
      v2 |= fkey(cz[i]) ^ k1;
    }
    gp.sync();
    const u64 Mc = gp.sum(loc);
    v2 = gp.orr(v2);
    const u64 S1 = (top_k > 0) ? Mc : S1_full;
    // bracket mode: the exact threshold is only known to lie in [thr_lo, thr_hi]; select with thr_hi and
    // report whether every threshold of the bracket gives the same kept set (selection is monotone in thr)
    const u64 thr = ambiguous ? thr_hi : scale_q32(S1, tpq);
    if (ambiguous) *ambiguous = (Mc > thr_hi) ? 0 : 1;
    cutkey = group_select_key<true, BLK>(gp, cz, cw, m, k1 & ~v2, v2, thr - G_off);
    u64 g = 0, cnt = 0;
    for (int i = gp.first(); i < m; i += gp.stride()) {
      const unsigned key = fkey(cz[i]);
      if (key > cutkey) g += cw[i];
      if (key == cutkey) ++cnt;
    }
    const u64 Gc = G_off + gp.sum(g), cntc = gp.sum(cnt);
    const u64 wc = fix40(cweight(fkey_inv(cutkey), c1, mc1));
    u64 mkeep = cntc;
    if (wc > 0 && thr >= Gc) {
      const u64 q = (thr - Gc) / wc + 1ull;
      mkeep = q < cntc ? q : cntc;
    }
    if (ambiguous) {
      u64 mk_lo = cntc;
      if (Gc > thr_lo) *ambiguous = 1;
      else if (wc > 0) { const u64 q = (thr_lo - Gc) / wc + 1ull; mk_lo = q < cntc ? q : cntc; }
      if (mk_lo != mkeep) *ambiguous = 1;
    }
    if (mkeep < cntc) {  // keep the first mkeep ties in ascending index order
      int lo = 0, hi = V - 1;
      if (BLK && cntc <= (u64)blockDim.x) {
        // tie group fits one element per thread: gather its indices (the masses in cw are no longer needed),
        // then every bisection step is a single hardware barrier-count
        int* tl = (int*)cw;
        __syncthreads();
        if (threadIdx.x == 0) gp.shu[0] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x)
          if (fkey(cz[i]) == cutkey) tl[atomicAdd(&gp.shu[0], 1u)] = cj[i];
        __syncthreads();
        const int myj = ((u64)threadIdx.x < cntc) ? tl[threadIdx.x] : 0x7FFFFFFF;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((u64)__syncthreads_count(myj <= mid) >= mkeep) hi = mid; else lo = mid + 1;
        }
      } else {
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          u64 c2 = 0;
          for (int i = gp.first(); i < m; i += gp.stride()) c2 += (fkey(cz[i]) == cutkey && cj[i] <= mid) ? 1u : 0u;
          c2 = gp.sum(c2);
          if (c2 >= mkeep) hi = mid; else lo = mid + 1;
        }
      }
      jcut = lo;
    }
  }
  u64 loc = 0;
  for (int i = gp.first(); i < m; i += gp.stride()) {
    const unsigned key = fkey(cz[i]);
    if (key > cutkey || (key == cutkey && cj[i] <= jcut)) loc += fix40(cweight(cz[i], c, mc));
  }
  Sfix_out = S_above + gp.sum(loc);
  cut_out = fkey_inv(cutkey);
  jcut_out = jcut;
  if (m_out) *m_out = m;  // cz / cj [0, m) = the candidates that survived top-k (all of them without top-k)
}

// Warp-collective: the kept tokens among candidates [0, m) -> dst (unordered); returns their number, 0 if more than KL_MAX.
__device__ __forceinline__ int write_kept_list(const float* cz, const int* cj, int m, float cut, int jcut, int2* dst) {
  const int lane = threadIdx.x & 31;
  int cnt = 0;
  for (int base = 0; base < m; base += 32) {
    const int i = base + lane;
    const float z = (i < m) ? cz[i] : 0.0f;
    const int j = (i < m) ? cj[i] : 0;
    const bool keep = (i < m) && (z > cut || (z == cut && j <= jcut));
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < KL_MAX) dst[pos] = make_int2(__float_as_int(z), j);
    cnt += __popc(bal);
  }
  return cnt <= KL_MAX ? cnt : 0;
}

// ---------------------------------------------------------------------------------------------
// rowstats_kernel
// ---------------------------------------------------------------------------------------------
template <int DT, bool HK, bool HP>
__global__ void __launch_bounds__(RS_NT, 2) rowstats_kernel(RowJob job_in) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // PDL: plan_kernel may be scheduled early (it waits)
  // specialised per mask type so that each instantiation only carries its own code path
  RowJob job = job_in;
  if (!HK) job.top_k = 0;
  if (!HP) job.use_p = 0;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ unsigned shu[33];
  __shared__ int s_count;
  __shared__ float s_cut;
  __shared__ int s_jcut;
  __shared__ u64 s_Sfix;
  const int V = job.V;
  const float c = job.c;
  constexpr bool masked = HK || HP;
  if (job.skip_resolved && job.n_unres && __ldcg(job.n_unres) == 0) return;  // rowsel_tma_kernel resolved every row
  for (long long r = blockIdx.x; r < job.R; r += gridDim.x) {
    if (job.skip_resolved && (job.out[r].flags & 1)) continue;  // done by rowsel_tma_kernel / nucleus_fast_kernel (block-uniform)
    const void* row = row_ptr<DT>(job, r);
    const bool aligned = (((size_t)row) & 15) == 0;
    // sweep 1: row max (+ per-thread maxima, reused as candidate thresholds)
    float tmax = -INFINITY;
    sweep<DT>(row, V, aligned, [&](const float(&x)[8], int) {
#pragma unroll
      for (int k = 0; k < 8; ++k) tmax = fmaxf(tmax, x[k]);
    });
    const float m = block_max_f(tmax, shf);
    const float mc = __fmul_rn(m, c);
    float cut = -INFINITY;
    int jcut = V, kcnt = 0;
    u64 Sfix = 0;
    if (!masked) {
      u64 s = 0;
      sweep<DT>(row, V, aligned, [&](const float(&x)[8], int) { s += sum_fix40_8(x, c, mc); });
      Sfix = block_sum_u64(s, sh64);
    } else {
      const float c1 = job.c1;
      const float mc1 = __fmul_rn(m, c1);
      // thresholds guaranteeing >= 16 / 64 / 256 / 512 elements above them (RS_NT = 512 threads)
      float pm = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
      float qm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, 2));
      qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, 4));
      float wm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, 8));
      wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, 16));
      float tau[4];
      tau[0] = block_min_f(wm, shf);
      tau[1] = block_min_f(qm, shf);
      tau[2] = block_min_f(pm, shf);
      tau[3] = block_min_f(tmax, shf);
      int L = -1;
      u64 S1 = 0;
      u64 band_G = 0;
      if (HK) {
        if (job.top_k <= RS_NT) {
          // threshold = k-th largest of the per-thread maxima: >= k elements are guaranteed above it and
          // only ~k(1+k/2T) are expected.  Bit-wise bisection, one hardware barrier-count per varying bit.
          const unsigned mykey = fkey(tmax);
          const unsigned refkey = fkey(m);
          unsigned vary = __reduce_or_sync(0xffffffffu, mykey ^ refkey);
          if ((threadIdx.x & 31) == 0) shu[threadIdx.x >> 5] = vary;
          __syncthreads();
          vary = 0;
#pragma unroll
          for (int w = 0; w < RS_NT / 32; ++w) vary |= shu[w];
          __syncthreads();
          unsigned K = refkey & ~vary;
          for (int bit = 31; bit >= 0; --bit) {
            if (!((vary >> bit) & 1u)) continue;
            const unsigned tr = K | (1u << bit);
            if (__syncthreads_count(mykey >= tr) >= job.top_k) K = tr;
          }
          tau[0] = fkey_inv(K);
          L = 0;
        } else {
          L = -1;
        }
      }
      float* cz = (float*)dyn_smem;
      int* cj = (int*)(dyn_smem + (size_t)CAP * 4);
      u64* cw = (u64*)(dyn_smem + (size_t)CAP * 8);
      // Compaction of the elements in [th, up) into shared memory (warp-aggregated slot allocation, one vote
      // per vector gates the per-element work, 4 loads in flight).  sa != nullptr additionally accumulates
      // the exact T=1 mass of EVERY element (pure nucleus, first pass).  Returns the number of elements
      // in the range (which may exceed CAP: then the buffer content is incomplete).
      auto collect = [&](const float th, const float up, u64* sa) -> int {
        const int NVr = (V + 7) >> 3, lane = threadIdx.x & 31;
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        auto emit = [&](const float(&x)[8], int v) {
          if (sa) *sa += sum_fix40_8(x, c1, mc1);  // (-inf padding contributes 0)
          compact_vec(x, v, th, up, V, cz, cj, &s_count, CAP);
        };
        for (int base = (threadIdx.x >> 5) << 5; base < NVr; base += 4 * RS_NT) {  // warp-uniform
          float x0[8], x1[8], x2[8], x3[8];
          const int v0 = base + lane, v1 = v0 + RS_NT, v2 = v0 + 2 * RS_NT, v3 = v0 + 3 * RS_NT;
          load8<DT>(row, min(v0, NVr - 1), V, aligned, x0);
          load8<DT>(row, min(v1, NVr - 1), V, aligned, x1);
          load8<DT>(row, min(v2, NVr - 1), V, aligned, x2);
          load8<DT>(row, min(v3, NVr - 1), V, aligned, x3);
          if (v0 >= NVr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x0[k] = -INFINITY;
          }
          if (v1 >= NVr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x1[k] = -INFINITY;
          }
          if (v2 >= NVr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x2[k] = -INFINITY;
          }
          if (v3 >= NVr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x3[k] = -INFINITY;
          }
          emit(x0, v0);
          if (base + RS_NT < NVr) emit(x1, v1);
          if (base + 2 * RS_NT < NVr) emit(x2, v2);
          if (base + 3 * RS_NT < NVr) emit(x3, v3);
        }
        __syncthreads();
        const int cnt = s_count;
        __syncthreads();
        return cnt;
      };
      int n = 0;
      u64 S_above = 0;
      if (HK) {
        if (L >= 0) {
          // Candidates (z >= th) can only sit in the slices of threads whose own maximum is >= th -- about k of
          // the 512 threads.  Instead of a second sweep over the whole row, each warp re-reads just those
          // slices (vector v = t + l*RS_NT of thread t: one gather per 32 vectors) and compacts the hits.
          const float th = tau[L];
          const int NVr = (V + 7) >> 3, lane = threadIdx.x & 31;
          if (threadIdx.x == 0) s_count = 0;
          __syncthreads();
          unsigned bal = __ballot_sync(0xffffffffu, tmax >= th);
          while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const int tbase = (threadIdx.x & ~31) + src;
            for (int l0 = 0; tbase + l0 * RS_NT < NVr; l0 += 32) {  // warp-uniform
              const int v = tbase + (l0 + lane) * RS_NT;
              float x[8];
              load8<DT>(row, min(v, NVr - 1), V, aligned, x);
              if (v >= NVr) {
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = -INFINITY;
              }
              compact_vec(x, v, th, INFINITY, V, cz, cj, &s_count, CAP);
            }
          }
          __syncthreads();
          n = s_count;
          __syncthreads();
          if (n > CAP || n < job.top_k || n == 0) L = -1;
        }
      } else {
        // pure nucleus, exact: pass A = exact T=1 mass of the row merged with the compaction of everything
        // above the thread-maxima threshold; then bands of ~CAP/2 elements are taken downwards (one cheap
        // compaction sweep each) until the exact cumulative mass crosses thr.  Every consumed band is kept
        // entirely, so its exact masses are simply added up.
        u64 sa_local = 0;
        float lo = tau[3], hi = INFINITY;
        n = collect(lo, hi, &sa_local);
        S1 = block_sum_u64(sa_local, sh64);
        const u64 thr = scale_q32(S1, job.tpq);
        u64 G_hi = 0;
        // first band below tau3: ~55 % of the n elements >= tau3 lie in [tau3, tau2) (order statistics of the
        // thread / pair maxima), which gives the local density; aim at CAP/2 elements with a 0.75 safety factor
        float width = (m - tau[3]) * 0.125f;
        if (tau[2] > tau[3] && n > 0) width = (tau[2] - tau[3]) * ((float)(CAP / 2) / (0.55f * (float)n)) * 0.75f;
        for (int iter = 0; iter < 40; ++iter) {
          if (n > CAP) {  // too many elements in [lo, hi): halve the band by value
            const float top = fminf(hi, m);
            const float mid = lo + (top - lo) * 0.5f;
            if (!(lo > -INFINITY) || !(mid > lo) || !(mid < top)) break;  // a tie group larger than CAP: slow path
            lo = mid;
          } else {
            u64 loc = 0, locT = 0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
              loc += fix40(cweight(cz[i], c1, mc1));
              if (c != c1) locT += fix40(cweight(cz[i], c, mc));
            }
            const u64 Mc = block_sum_u64(loc, sh64);
            if (G_hi + Mc > thr) { band_G = G_hi; L = 0; break; }  // the cut lies in this band
            S_above += (c != c1) ? block_sum_u64(locT, sh64) : Mc;
            G_hi += Mc;
            if (!(lo > -INFINITY)) break;
            if (n > 0 && hi < INFINITY) width = (hi - lo) * ((float)(CAP / 2) / (float)n) * 0.75f;  // density grows downwards
            if (!(width > 0.0f)) width = 1.0f;
            hi = lo;
            lo = (iter >= 30) ? -INFINITY : hi - width;
          }
          n = collect(lo, hi, nullptr);
        }
        if (n == 0 || n > CAP) L = -1;
      }
      if (L >= 0 && n <= WARP_SELECT_MAX) {
        if (threadIdx.x < 32) {
          const Grp<false> gp{sh64, shu};
          int mval = 0, kc = 0;
          select_cut_group<false>(gp, cz, cj, cw, n, V, job.top_k, job.use_p, job.tpq, S1, band_G, S_above, c, mc, c1, mc1,
                                  cut, jcut, Sfix, 0, 0, nullptr, &mval);
          // with top-k the candidates are all elements above a threshold, i.e. a superset of the kept set: small kept
          // sets are handed to sample_lists_kernel as a list, which then never sweeps the row again
          if (HK && job.klist) kc = write_kept_list(cz, cj, mval, cut, jcut, job.klist + r * KL_MAX);
          if (threadIdx.x == 0) { s_cut = cut; s_jcut = jcut; s_Sfix = Sfix; s_count = kc; }
        }
        __syncthreads();
        cut = s_cut; jcut = s_jcut; Sfix = s_Sfix; kcnt = s_count;
        __syncthreads();
      } else if (L >= 0 && job.top_k == 0) {
        const Grp<true> gp{sh64, shu};
        select_cut_group<true>(gp, cz, cj, cw, n, V, 0, job.use_p, job.tpq, S1, band_G, S_above, c, mc, c1, mc1, cut, jcut,
                               Sfix);
      } else if (L >= 0) {
        CandSrc src{cz, cj, cw, n, c1, mc1};
        select_cut(src, V, job.top_k, job.use_p, job.tpq, S1, c, mc, sh64, shu, cut, jcut, Sfix);
      } else {  // exact but slow: every selection step is a sweep over the row
        RowSrc<DT> src{row, V, aligned, c1, mc1};
        if (job.use_p && job.top_k == 0 && S1 == 0) {
          u64 sa = 0;
          src.each(0u, true, [&](float, unsigned, int, u64 w) { sa += w; });
          S1 = block_sum_u64(sa, sh64);
        }
        select_cut(src, V, job.top_k, job.use_p, job.tpq, S1, c, mc, sh64, shu, cut, jcut, Sfix);
      }
    }
    if (threadIdx.x == 0) {
      RowOut o;
      o.m = m; o.mc = mc;
      const float S32 = __fmul_rn(__ull2float_rn(Sfix), 0x1p-40f);
      o.inv = __fdiv_rn(1.0f, S32);
      o.cut = cut; o.jcut = jcut; o.flags = 1 | (kcnt << 8); o.Sfix = Sfix;
      job.out[r] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// nucleus_fast_kernel (pure top-p): resolves rows whose nucleus is small -- i.e. real LLM logits --
// without an exact pass over the vocabulary.  The cut only needs (a) the exact T=1 masses of the few
// top tokens and (b) thr = top_p * S1, and S1 is only needed to ~1e-4 unless a prefix mass happens to
// lie that close to thr.  So: sweep 1 = max + online MUFU sum (S1 to 1e-6); sweep 2 = compaction of
// the candidates above the 64-group threshold + MUFU masses above the nested thresholds; exact masses
// of the candidates; selection with the bracket [thr(1-1e-4), thr(1+1e-4)].  If both ends of the bracket
// give the same kept set the result equals the exact one (selection is monotone in thr) and the row is
// flagged resolved; otherwise rowstats_kernel<.., false, true> redoes it exactly.
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(RS_NT, 2) nucleus_fast_kernel(RowJob job) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ unsigned shu[33];
  __shared__ int s_count, s_amb, s_kc;
  __shared__ float s_cut;
  __shared__ int s_jcut;
  __shared__ u64 s_Sfix;
  float* cz = (float*)dyn_smem;
  int* cj = (int*)(dyn_smem + (size_t)CAP * 4);
  u64* cw = (u64*)(dyn_smem + (size_t)CAP * 8);
  const int V = job.V, NV = (V + 7) >> 3, lane = threadIdx.x & 31;
  const float c = job.c, c1 = job.c1;
  for (long long r = blockIdx.x; r < job.R; r += gridDim.x) {
    const void* row = row_ptr<DT>(job, r);
    const bool aligned = (((size_t)row) & 15) == 0;
    float m, S1f, th, Wt;
    if (job.pre_stats) {  // block-uniform: everything sweep 1 produces was emitted by rowfast_tma_kernel<DT, 1>
      const RowOut pre = job.out[r];
      m = pre.m; S1f = pre.inv; th = pre.cut; Wt = __uint_as_float((unsigned)pre.Sfix);
      __syncthreads();  // every thread has read the record before thread 0 overwrites it below
    } else {
      // sweep 1: exact max, per-thread maxima, online MUFU sum at T = 1
      float tm = -INFINITY, ts = 0.0f;
      sweep<DT>(row, V, aligned, [&](const float(&x)[8], int) {
        const float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
        if (vm > tm) { ts = __fmul_rn(ts, ex2_approx(__fmul_rn(__fsub_rn(tm, vm), c1))); tm = vm; }
        const float mcl = (tm > -INFINITY) ? __fmul_rn(tm, c1) : 0.0f;  // (only -inf so far: no NaN)
#pragma unroll
        for (int k = 0; k < 8; ++k) ts = __fadd_rn(ts, ex2_approx(__fmaf_rn(x[k], c1, -mcl)));
      });
      m = block_max_f(tm, shf);
      const float resc = (tm > -INFINITY) ? ex2_approx(__fmul_rn(__fsub_rn(tm, m), c1)) : 0.0f;  // weight of my maximum
      ts = __fmul_rn(ts, resc);
      float wt = resc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ts += __shfl_xor_sync(0xffffffffu, ts, o);
        wt += __shfl_xor_sync(0xffffffffu, wt, o);
      }
      if (lane == 0) { shf[threadIdx.x >> 5] = ts; shf[16 + (threadIdx.x >> 5)] = wt; }
      __syncthreads();
      S1f = 0.0f; Wt = 0.0f;
#pragma unroll
      for (int w = 0; w < RS_NT / 32; ++w) { S1f += shf[w]; Wt += shf[16 + w]; }
      __syncthreads();
      // candidate threshold: min over the 8-lane-group maxima (>= RS_NT/8 elements above it)
      float qm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
      qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, 2));
      qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, 4));
      th = block_min_f(qm, shf);
    }
    const float mc = __fmul_rn(m, c), mc1 = __fmul_rn(m, c1);
    // Flat rows cannot be resolved here (their nucleus holds thousands of tokens): if the per-thread maxima --
    // roughly the 512 (256 with the TMA pre-pass) largest logits -- carry less than 3/4 of the mass the nucleus
    // needs, sweep 2 is skipped and the row goes to nucleus_hist_kernel.  (A heuristic: it only decides which
    // exact kernel does the row.)
    const bool flat = Wt < 0.75f * (float)((double)job.tpq * (1.0 / 4294967296.0)) * S1f;  // block-uniform
    // sweep 2: compaction of the candidates (warp-aggregated, vote-gated)
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int base = (threadIdx.x >> 5) << 5; base < NV && !flat; base += 4 * RS_NT) {
      float x[4][8];
      int vv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        vv[q] = base + q * RS_NT + lane;
        load8<DT>(row, min(vv[q], NV - 1), V, aligned, x[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (base + q * RS_NT >= NV) break;  // warp-uniform
        const int v = vv[q];
        if (v >= NV) {
#pragma unroll
          for (int k = 0; k < 8; ++k) x[q][k] = -INFINITY;
        }
        compact_vec(x[q], v, th, INFINITY, V, cz, cj, &s_count, CAP);
      }
    }
    __syncthreads();
    const int n = s_count;
    bool ok = (n > 0) && (n <= WARP_SELECT_MAX);
    if (ok) {
      if (threadIdx.x < 32) {
        // S1 in 2^-40 fixed point, bracketed by the MUFU error (1e-4 is ~50x the observed 2e-6)
        const double S1d = (double)S1f * 1099511627776.0;
        const u64 thr_lo = scale_q32((u64)(S1d * (1.0 - 1e-4)), job.tpq), thr_hi = scale_q32((u64)(S1d * (1.0 + 1e-4)), job.tpq);
        const Grp<false> gp{sh64, shu};
        float cut; int jcut; u64 Sfix; int amb = 1;
        select_cut_group<false>(gp, cz, cj, cw, n, V, 0, 1, job.tpq, 0ull, 0ull, 0ull, c, mc, c1, mc1, cut, jcut, Sfix,
                                thr_lo, thr_hi, &amb);
        int kc = 0;  // (unambiguous: the candidates hold the whole nucleus)
        if (!amb && job.klist) kc = write_kept_list(cz, cj, n, cut, jcut, job.klist + r * KL_MAX);
        if (lane == 0) { s_cut = cut; s_jcut = jcut; s_Sfix = Sfix; s_amb = amb; s_kc = kc; }
      }
      __syncthreads();
      ok = (s_amb == 0);
    }
    if (threadIdx.x == 0) {
      RowOut o;
      o.m = m; o.mc = mc; o.cut = -INFINITY; o.jcut = V; o.flags = 0; o.Sfix = 0;
      o.inv = S1f;  // unresolved rows: the MUFU T=1 mass (relative to the max) for nucleus_hist_kernel
      if (ok) {
        o.cut = s_cut; o.jcut = s_jcut; o.Sfix = s_Sfix; o.flags = 1 | (s_kc << 8);
        o.inv = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(s_Sfix), 0x1p-40f));
      }
      job.out[r] = o;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// nucleus_hist_kernel (pure top-p, rows nucleus_fast_kernel left unresolved = flat rows whose nucleus holds
// thousands of tokens, up to ~90 % of the vocabulary for random-init models).  The cut is located by a
// radix-select over the VALUE axis whose cost does not depend on the size of the nucleus:
//   level sweeps (approximate, MUFU): every thread counts its elements into a PRIVATE 256-bin byte
//     histogram in shared memory (column tid of a [256][512] byte array: no atomics, no bank conflicts),
//     bins uniform in (top - z); bin masses = count x mass(bin centre), with the rigorous half-bin bound
//     as slack; the bins that are certainly kept / possibly kept bracket the cut.  The next level zooms
//     into the bracket (256x finer) until it holds <= NH_CAP elements (1 level for 3*randn rows, 2-3 for
//     near-uniform rows).
//   exact sweep: canonical T=1 mass of every element (S1), exact mass / tempered weight of everything
//     above the bracket, compaction of the bracket's elements; the cut is then selected among them by
//     select_cut_group with exact integers.  The bracket is VERIFIED with the exact sums
//     (G_above <= thr < G_above + M_band); a row failing any check stays unresolved and is redone by
//     rowstats_kernel, so the histogram levels can only cost time, never change a result.
// ---------------------------------------------------------------------------------------------
constexpr int NH_T = 512, NH_BINS = 256, NH_CAP = 8192, NH_RV = 32, NH_SLOTS = 32;  // private candidate slots per thread (indices only)
// failure statistics of nucleus_hist_kernel (attempts / rows): [0] inconsistent estimate, [1] byte counter wrapped,
// [2] nucleus reaches the 2^-40 tail, [3] levels exhausted, [4] private slots overflowed, [5] bracket above the cut,
// [6] bracket below the cut, [7] rows left to the slow path, [8] rows resolved, [9] attempts
__device__ unsigned long long g_nh_stats[16];
constexpr size_t NH_SMEM = (size_t)(NH_BINS + 1) * NH_T;  // 128.5 KB (256 bins + a spare); later reused for the candidates (NH_CAP * 16 B)

template <int DT>
__global__ void __launch_bounds__(NH_T, 1) nucleus_hist_kernel(RowJob job) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ unsigned cnt[NH_BINS];
  __shared__ float wsum[NH_T / 32];
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ unsigned shu[33];
  __shared__ int s_count;
  unsigned char* hist = dyn_smem;
  // exact stage: [0, 64 KB) private candidate slots, later the masses cw; [64 KB, 128 KB) dense candidates
  u64* cw = (u64*)dyn_smem;
  float* cz = (float*)(dyn_smem + (size_t)NH_CAP * 8);
  int* cj = (int*)(dyn_smem + (size_t)NH_CAP * 12);
  const int V = job.V, NV = (V + 7) >> 3, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float c = job.c, c1 = job.c1;
  const float top_p = (float)((double)job.tpq * (1.0 / 4294967296.0));
  // byte column of this thread inside a 512-byte bin row: the 32 lanes of a warp hit 32 different 4-byte words
  // (byte stores of four lanes into one word would be serialised)
  const int hcol = (tid & 127) * 4 + (tid >> 7);
  for (long long r = blockIdx.x; r < job.R; r += gridDim.x) {
    const RowOut ro = job.out[r];
    if (ro.flags & 1) continue;
    const float m = ro.m, S1f = ro.inv;
    if (!(m > -INFINITY) || !(m < INFINITY) || !(S1f > 0.0f) || !(S1f < INFINITY)) continue;  // block-uniform
    const void* row = row_ptr<DT>(job, r);
    const bool aligned = (((size_t)row) & 15) == 0;
    const float mc = __fmul_rn(m, c), mc1 = __fmul_rn(m, c1);
    const float thr_f = S1f * top_p;
    // A failed attempt (inconsistent estimates, bracket missed the cut, too many candidates) is retried with
    // 4x the slack and half the candidate budget before the row is left to the slow path.
    bool done = false;
    for (int attempt = 0; attempt < 3 && !done; ++attempt) {
    const float slack_mul = (float)(1 << (2 * attempt));
    const u64 focus_max = (u64)((NH_CAP / 2) >> attempt);
    // ---- histogram levels: bracket (zb, zt] of the value axis that contains the cut ----
    float zt = m, zb = -INFINITY;
    float scale = (float)NH_BINS / 27.725887f;  // level 0: x = m - z in [0, 40 ln 2): masses below 2^-40 are 0
    bool ok = false;
    for (int level = 0; level < 5 && !ok; ++level) {
      float fa = 0.0f;  // MUFU mass of the elements above zt
      for (int i = tid; i < NH_BINS; i += NH_T) cnt[i] = 0;
      for (int base = 0; base < NV; base += NH_RV * NH_T) {  // rounds of <= 256 elements per thread (byte counters)
        uint4* h4 = (uint4*)hist;
#pragma unroll
        for (int q = 0; q < NH_BINS / 16; ++q) h4[q * NH_T + tid] = make_uint4(0u, 0u, 0u, 0u);  // (spare bin: never read)
        __syncthreads();
        const int vend = min(NV, base + NH_RV * NH_T);
        sweep_range<DT, NH_T>(row, V, aligned, base, vend, [&](const float(&x)[8], int) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float z = x[k];
            if (level == 0) {
              const int b = min(__float2int_rz(__fmul_rn(__fsub_rn(zt, z), scale)), NH_BINS - 1);
              hist[b * NH_T + hcol]++;
            } else {  // branch-free: elements outside (zb, zt] are counted in the spare bin NH_BINS
              const float e = ex2_approx(__fmaf_rn(z, c1, -mc1));
              const float t = fminf(__fmul_rn(__fsub_rn(zt, z), scale), (float)NH_BINS);
              const bool above = z > zt;
              fa += above ? e : 0.0f;
              hist[__float2int_rz(above ? (float)NH_BINS : t) * NH_T + hcol]++;
            }
          }
        });
        __syncthreads();
        for (int b = warp; b < NH_BINS; b += NH_T / 32) {
          const unsigned* hr = (const unsigned*)(hist + b * NH_T);
          unsigned sacc = 0;
#pragma unroll
          for (int q = 0; q < NH_T / 128; ++q) sacc = __dp4a(hr[lane + 32 * q], 0x01010101u, sacc);
          sacc = __reduce_add_sync(0xffffffffu, sacc);
          if (lane == 0) cnt[b] += sacc;
        }
        __syncthreads();
      }
      // mass above the range (level 0: nothing is above the row maximum)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) fa += __shfl_xor_sync(0xffffffffu, fa, o);
      if (lane == 0) wsum[warp] = fa;
      __syncthreads();
      float A_above = 0.0f;
#pragma unroll
      for (int w = 0; w < NH_T / 32; ++w) A_above += wsum[w];
      __syncthreads();
      // estimated cumulative mass per bin (threads 0..255 = bins), half-bin worst case as slack
      // Slack of the estimate (the bracket is verified exactly afterwards, so it only has to be usually right):
      // statistical part = 4 sigma of the bin masses (cnt elements uniform over the bin, half-width `rel`),
      // systematic part (density slope inside a bin, convexity of exp) = 25 % of the half-bin bound, + MUFU / S1f error.
      const float rel = 0.36f * c1 / scale + 1e-6f;
      float massb = 0.0f, varb = 0.0f;
      unsigned myc = 0;
      if (tid < NH_BINS) {
        myc = cnt[tid];
        const float zmid = zt - ((float)tid + 0.5f) / scale;
        massb = (float)myc * ex2_approx(fminf(__fmaf_rn(zmid, c1, -mc1), 0.0f));
        varb = myc ? (rel * massb) * (rel * massb) / (3.0f * (float)myc) : 0.0f;
      }
      float incl = massb, vincl = varb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o), t2 = __shfl_up_sync(0xffffffffu, vincl, o);
        if (lane >= o) { incl += t; vincl += t2; }
      }
      if (lane == 31) { wsum[warp] = incl; shf[warp] = vincl; }
      __syncthreads();
      float off = 0.0f, voff = 0.0f;
      for (int w = 0; w < warp; ++w) { off += wsum[w]; voff += shf[w]; }
      const float cumE = off + incl, cumX = cumE - massb, varE = voff + vincl, varX = varE - varb;
      // (never more than the half-bin worst case rel * cum: rows whose top token dominates the mass)
      const float slE = slack_mul * fminf(rel * cumE, 5.0f * sqrtf(varE) + 0.25f * rel * cumE) + 1e-4f * (A_above + cumE);
      const float slX = slack_mul * fminf(rel * cumX, 5.0f * sqrtf(fmaxf(varX, 0.0f)) + 0.25f * rel * cumX) + 1e-4f * (A_above + cumX);
      const bool certain = tid < NH_BINS && (A_above + cumE + slE <= thr_f);
      const bool possible = tid < NH_BINS && (A_above + cumX - slX <= thr_f);
      const int nc = __syncthreads_count(certain), np = __syncthreads_count(possible);
      u64 tot_c = block_sum_u64((u64)myc, sh64);
      if (nc >= NH_BINS || np < 1) { if (tid == 0) atomicAdd(&g_nh_stats[0], 1ull); break; }  // inconsistent estimates
      if (level == 0 && tot_c != (u64)NV * 8ull) { if (tid == 0) atomicAdd(&g_nh_stats[1], 1ull); break; }  // a byte counter wrapped (giant tie group)
      const int b_lo = max(nc - 1, 0), b_hi = max(np - 1, b_lo);
      if (level == 0 && b_hi == NH_BINS - 1) { if (tid == 0) atomicAdd(&g_nh_stats[2], 1ull); break; }  // nucleus reaches the 2^-40 tail
      const u64 n_focus = block_sum_u64((tid >= b_lo && tid <= b_hi && tid < NH_BINS) ? (u64)myc : 0ull, sh64);
      const float nzt = (b_lo > 0) ? zt - (float)b_lo / scale : zt;
      const float nzb = zt - (float)(b_hi + 1) / scale;
      if (!(nzb < nzt)) break;
      zt = nzt; zb = nzb;
      scale = (float)NH_BINS / (zt - zb);
      if (n_focus <= focus_max) ok = n_focus > 0;  // first attempt: average 8 of the 32 private slots per thread
      else if (!(scale < 1e30f)) break;
    }
    if (tid == 0) atomicAdd(&g_nh_stats[9], 1ull);
    if (!ok) { if (tid == 0) atomicAdd(&g_nh_stats[3], 1ull); continue; }  // block-uniform
    // ---- exact sweep: S1, mass / tempered weight above the bracket, candidates (zb, zt] ----
    // Candidate indices go to PRIVATE slots (slot-major [NH_SLOTS][512] ints: no votes, scans or atomics in
    // the sweep); they are made dense afterwards (values re-read through L2).  A thread with more than
    // NH_SLOTS hits leaves the row to the slow path.
    const float th = nextafterf(zb, INFINITY), up = nextafterf(zt, INFINITY);  // (zb, zt] == [th, up), th > -inf
    int* stg = (int*)dyn_smem;
    int pc = 0;
    u64 S1l = 0, Gl = 0, STl = 0;
    const bool tempered = (c != c1);
    const float2 c12 = make_float2(c1, c1), nmc12 = make_float2(-mc1, -mc1), cT2 = make_float2(c, c), nmcT2 = make_float2(-mc, -mc);
    __syncthreads();
    for (int base = (tid >> 5) << 5; base < NV; base += 2 * NH_T) {  // warp-uniform
      float x0[8], x1[8];
      const int v0 = base + lane, v1 = v0 + NH_T;
      load8<DT>(row, min(v0, NV - 1), V, aligned, x0);
      load8<DT>(row, min(v1, NV - 1), V, aligned, x1);
      auto emit = [&](float(&x)[8], int v) {
        if (v >= NV) {
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 z2 = make_float2(x[2 * k], x[2 * k + 1]);
          const float2 e = cweight2(z2, c12, nmc12);
          const u64 w0 = fix40(e.x), w1 = fix40(e.y);
          S1l += w0 + w1;
          const bool a0 = z2.x >= up, a1 = z2.y >= up;
          if (a0) Gl += w0;
          if (a1) Gl += w1;
          if (tempered) {
            const float2 et = cweight2(z2, cT2, nmcT2);
            if (a0) STl += fix40(et.x);
            if (a1) STl += fix40(et.y);
          }
          if (!a0 && z2.x >= th) {  // (-inf padding never qualifies)
            if (pc < NH_SLOTS) stg[pc * NH_T + tid] = v * 8 + 2 * k;
            ++pc;
          }
          if (!a1 && z2.y >= th) {
            if (pc < NH_SLOTS) stg[pc * NH_T + tid] = v * 8 + 2 * k + 1;
            ++pc;
          }
        }
      };
      emit(x0, v0);
      if (base + NH_T < NV) emit(x1, v1);  // warp-uniform
    }
    // dense candidate list: block exclusive scan of the per-thread counts, then each thread moves its entries
    int incl_pc = pc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl_pc, o);
      if (lane >= o) incl_pc += t;
    }
    __syncthreads();
    if (lane == 31) shu[warp] = (unsigned)incl_pc;
    const int over = __syncthreads_count(pc > NH_SLOTS);
    int off_pc = incl_pc - pc, n = 0;
    for (int w = 0; w < NH_T / 32; ++w) { const int t = (int)shu[w]; if (w < warp) off_pc += t; n += t; }
    if (!over && n <= NH_CAP) {
      float* czd = (float*)(dyn_smem + (size_t)NH_CAP * 8);
      int* cjd = (int*)(dyn_smem + (size_t)NH_CAP * 12);
      for (int i = 0; i < pc; ++i) {
        const int j = stg[i * NH_T + tid];
        czd[off_pc + i] = load1<DT>(row, j);
        cjd[off_pc + i] = j;
      }
    }
    __syncthreads();
    const u64 S1 = block_sum_u64(S1l, sh64), G_above = block_sum_u64(Gl, sh64);
    const u64 S_above = tempered ? block_sum_u64(STl, sh64) : G_above;
    if (over || n <= 0 || n > NH_CAP) { if (tid == 0) atomicAdd(&g_nh_stats[4], 1ull); continue; }
    const u64 thr = scale_q32(S1, job.tpq);
    u64 mb = 0;
    for (int i = tid; i < n; i += NH_T) mb += fix40(cweight(cz[i], c1, mc1));
    const u64 M_band = block_sum_u64(mb, sh64);
    if (G_above > thr || G_above + M_band <= thr) {  // bracket missed the cut: retry wider
      if (tid == 0) atomicAdd(&g_nh_stats[G_above > thr ? 5 : 6], 1ull);
      continue;
    }
    float cut; int jcut; u64 Sfix;
    const Grp<true> gp{sh64, shu};
    select_cut_group<true>(gp, cz, cj, cw, n, V, 0, 1, job.tpq, S1, G_above, S_above, c, mc, c1, mc1, cut, jcut, Sfix);
    if (tid == 0) {
      RowOut o;
      o.m = m; o.mc = mc;
      o.inv = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(Sfix), 0x1p-40f));
      o.cut = cut; o.jcut = jcut; o.flags = 1; o.Sfix = Sfix;
      job.out[r] = o;
    }
    done = true;
    __syncthreads();
    }  // attempt
    if (tid == 0) atomicAdd(&g_nh_stats[done ? 8 : 7], 1ull);
  }
}

// ---------------------------------------------------------------------------------------------
// sampling helpers (block-wide)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool kept(const RowOut& ro, float z, int j) {
  return z > ro.cut || (z == ro.cut && j <= ro.jcut);
}
template <int DT>
__device__ __forceinline__ float row_prob(const RowOut& ro, const void* row, int j, float c) {
  const float z = load1<DT>(row, j);
  const float e = kept(ro, z, j) ? cweight(z, c, ro.mc) : 0.0f;
  return __fmul_rn(e, ro.inv);
}

// bit 63 of a u64 the megakernel exchanges between CTAs (partial sums < 2^62): set = written in this launch
constexpr u64 WORD_VALID = 1ull << 63;

// Given the per-256-element partial sums of integer weights, find the first index whose inclusive
// prefix sum exceeds target (warp 0 scans the partials, then re-evaluates one segment).  Block-wide
// call; returns -1 if target >= total.
template <typename WF>
__device__ long long locate_token(int NV, int P, const u64* part, u64 total, u64 target, WF wf, long long* s_res) {
  const int lane = threadIdx.x & 31;
  const bool pglobal = __isGlobal(part);  // (decide_kernel keeps its partials in shared memory)
  if (threadIdx.x < 32) {
    long long res = -1;
    if (target < total) {
      // coalesced scan: 32 partials per step, all loads of a step independent
      int seg = -1;
      u64 before = 0, run = 0;
      for (int base = 0; base < P && seg < 0; base += 128) {
        u64 pv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int i = base + q * 32 + lane; pv[q] = (i < P) ? ((pglobal ? __ldcg(&part[i]) : part[i]) & ~WORD_VALID) : 0ull; }  // (partials of other CTAs: L2-coherent read; the megakernel tags them valid)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          u64 incl = pv[q];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, run + incl > target);
          if (seg < 0 && bal) {
            const int owner = __ffs(bal) - 1;
            seg = base + q * 32 + owner;
            before = run + __shfl_sync(0xffffffffu, incl - pv[q], owner);
          }
          run += __shfl_sync(0xffffffffu, incl, 31);
        }
      }
      if (seg < 0) seg = 0;
      const int v = seg * 32 + lane;
      u64 w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = 0;
      if (v < NV) wf(v, w);
      u64 ls = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) ls += w[k];
      u64 li = ls;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, li, o);
        if (lane >= o) li += t;
      }
      const u64 loc_t = target - before;
      const unsigned b2 = __ballot_sync(0xffffffffu, li > loc_t);
      const int own2 = __ffs(b2) - 1;
      if (lane == own2) {
        u64 run = li - ls;
        int k = 0;
#pragma unroll
        for (; k < 8; ++k) {
          if (run + w[k] > loc_t) break;
          run += w[k];
        }
        res = (long long)v * 8 + k;
      }
      res = __shfl_sync(0xffffffffu, res, own2 < 0 ? 0 : own2);
    }
    if (lane == 0) *s_res = res;
  }
  cta_sync();
  const long long out = *s_res;
  cta_sync();
  return out;
}

// Inverse CDF over integer weights wf(v, w[8]) in index order: returns the first index whose
// inclusive prefix sum exceeds target.  part[] receives the per-256-element partial sums.
// total_out = sum of all weights.  If target >= total the result is -1.
template <typename WF>
__device__ long long invcdf_sweep(int V, WF wf, bool have_target, u64 target_in, unsigned u24, u64* part, u64* sh64,
                                  long long* s_res, u64& total_out) {
  const int NV = (V + 7) >> 3;
  const int P = (NV + 31) >> 5;
  const int lane = threadIdx.x & 31;
  for (int base = (threadIdx.x >> 5) << 5; base < NV; base += cta_nthreads()) {  // warp-uniform
    const int v = base + lane;
    u64 s = 0;
    if (v < NV) {
      u64 w[8];
      wf(v, w);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += w[k];
    }
    s = warp_sum_u64(s);
    if (lane == 0) part[base >> 5] = s;
  }
  cta_sync();
  u64 loc = 0;
  for (int i = threadIdx.x; i < P; i += cta_nthreads()) loc += part[i];
  const u64 total = block_sum_u64(loc, sh64);
  total_out = total;
  const u64 target = have_target ? target_in : scale_u24(total, u24);
  return locate_token(NV, P, part, total, target, wf, s_res);
}

// block-wide argmax, first index on ties; value must be >= 0; returns -1 if no value > floor_excl
__device__ long long block_argmax(float best, int idx, float* shf, int* shi, long long* s_res) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { shf[w] = best; shi[w] = idx; }
  cta_sync();
  if (threadIdx.x < 32) {
    const int nwarp = (cta_nthreads() + 31) >> 5;
    best = (lane < nwarp) ? shf[lane] : -INFINITY;
    idx = (lane < nwarp) ? shi[lane] : 0x7FFFFFFF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (lane == 0) *s_res = (idx == 0x7FFFFFFF) ? -1ll : (long long)idx;
  }
  cta_sync();
  const long long out = *s_res;
  cta_sync();
  return out;
}

struct Scratch {
  u64* part;
  u64* sh64;
  float* shf;
  int* shi;
  long long* s_res;
};

// sample() on a processed target row (utils/logits_processor.py:36 greedy / inverse CDF)
template <int DT>
__device__ long long sample_p_row(const void* row, const RowOut& ro, int V, float c, bool greedy, float u,
                                  const Scratch& sc) {
  const bool aligned = (((size_t)row) & 15) == 0;
  if (greedy) {
    float best = -1.0f;
    int idx = 0x7FFFFFFF;
    sweep<DT>(row, V, aligned, [&](const float(&x)[8], int j0) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (j0 + k < V) {
          const float e = kept(ro, x[k], j0 + k) ? cweight(x[k], c, ro.mc) : 0.0f;
          if (e > best) { best = e; idx = j0 + k; }
        }
    });
    return block_argmax(best, idx, sc.shf, sc.shi, sc.s_res);
  }
  u64 total;
  auto wf = [&](int v, u64(&w)[8]) {
    float x[8];
    load8<DT>(row, v, V, aligned, x);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      w[k] = (v * 8 + k < V && kept(ro, x[k], v * 8 + k)) ? fix40(cweight(x[k], c, ro.mc)) : 0ull;
  };
  return invcdf_sweep(V, wf, true, scale_u24(ro.Sfix, u24_of(u)), 0u, sc.part, sc.sh64, sc.s_res, total);
}

// max_fn(p - q) then sample (sampling/speculative_decoding.py:10-19,168,171); -1 => caller falls back to p
template <int DT>
__device__ long long sample_residual(const void* prow, const RowOut& rp, const void* qrow, const RowOut& rq, int V,
                                     float c, bool greedy, float u, u64 rmin, const Scratch& sc) {
  const bool pal = (((size_t)prow) & 15) == 0, qal = (((size_t)qrow) & 15) == 0;
  auto resid = [&](float zp, float zq, int j) -> float {
    const float P = __fmul_rn(kept(rp, zp, j) ? cweight(zp, c, rp.mc) : 0.0f, rp.inv);
    const float Q = __fmul_rn(kept(rq, zq, j) ? cweight(zq, c, rq.mc) : 0.0f, rq.inv);
    const float r = __fsub_rn(P, Q);
    return r > 0.0f ? r : 0.0f;
  };
  u64 total;
  float best = 0.0f;
  int idx = 0x7FFFFFFF;
  auto wf = [&](int v, u64(&w)[8]) {
    float xp[8], xq[8];
    load8<DT>(prow, v, V, pal, xp);
    load8<DT>(qrow, v, V, qal, xq);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = v * 8 + k;
      const float r = (j < V) ? resid(xp[k], xq[k], j) : 0.0f;
      w[k] = fix60(r);
      if (r > best || (r == best && r > 0.0f && j < idx)) { best = r; idx = j; }
    }
  };
  // u24 path: target depends on the total, which the sweep itself produces
  long long x = invcdf_sweep(V, wf, false, 0ull, u24_of(u), sc.part, sc.sh64, sc.s_res, total);
  if (total <= rmin) return -1;
  if (greedy) {
    const long long g = block_argmax(best, idx, sc.shf, sc.shi, sc.s_res);
    return g;
  }
  return x;
}

// Stop-token scan over the n accepted drafts.  sampling/speculative_decoding.py:150-152 (and ngram_assisted.py:124-126)
// take torch.nonzero(eq(ids[1,n], stop[k,1]))[0,1]: rows of the [k,n] match matrix come first, i.e. the FIRST-LISTED
// stop token that occurs anywhere among the accepted drafts decides, at its first position.  The batched engine
// (engine/infer_engine.py:310-312, SPECDEC_ACCEPT_BATCHED) stops at the earliest accepted position holding any end token.
__device__ __forceinline__ int first_stop_index(const long long* toks, int n, const long long* stop, int n_stop, int flags) {
  if (flags & SPECDEC_ACCEPT_BATCHED) {
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n_stop; ++k)
        if (toks[i] == stop[k]) return i;
    return -1;
  }
  for (int k = 0; k < n_stop; ++k)
    for (int i = 0; i < n; ++i)
      if (toks[i] == stop[k]) return i;
  return -1;
}

// ---------------------------------------------------------------------------------------------
// decide_kernel: one CTA per sequence
// ---------------------------------------------------------------------------------------------
struct DecideJob {
  RowJob rj;
  const long long* draft_tokens;
  const float* u_accept;
  const float* u_sample;
  u64 seed, offset;
  const u64* offset_dev;  // SPECDEC_OFFSET_DEVICE: the offset lives in device memory and is read when the kernels RUN
  long long seq0;
  int gamma;
  int greedy;
  int flags;
  const long long* stop;
  int n_stop;
  int* n_acc;
  long long* next_tok;
  unsigned char* mask;
  float* p_tok;
  float* q_tok;
  int* first_stop;
  float* next_prob;
  int* packed;
  int lane_sample;  // philox lane of the sample stream
};

// Philox offset of the step: a host scalar, or (SPECDEC_OFFSET_DEVICE) a word in device memory, so that a replayed
// CUDA graph / a device-resident decode loop advances its uniforms without new host arguments.
__device__ __forceinline__ u64 job_offset(const DecideJob& job) { return job.offset_dev ? __ldg(job.offset_dev) : job.offset; }

template <int DT>
__global__ void __launch_bounds__(NT, 1) decide_kernel(DecideJob job) {
  __shared__ u64 part[MAXPART];
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ int shi[33];
  __shared__ long long s_res;
  __shared__ int s_acc[64];
  __shared__ int s_n;
  const Scratch sc{part, sh64, shf, shi, &s_res};
  const RowJob& rj = job.rj;
  const int b = blockIdx.x, g = job.gamma, V = rj.V;
  const int rps = rj.nT + rj.nD;
  const float c = rj.c;
  const bool greedy = job.greedy != 0;
  const bool ngram = (job.flags & SPECDEC_NGRAM) != 0;
  const RowOut* ro = rj.out + (long long)b * rps;
  const unsigned seq = (unsigned)(job.seq0 + b);
  const long long* toks = job.draft_tokens + (long long)b * g;

  // 1. per-position accept test
  if (ngram) {
    for (int i = 0; i < g; ++i) {  // accept iff draft == sample(p_i)  (ngram_assisted.py:114-119)
      const float u = job.u_accept ? job.u_accept[(long long)b * g + i] : philox_uniform(job.seed, job_offset(job), seq, i);
      const void* prow = row_ptr<DT>(rj, (long long)b * rps + i);
      const long long s = sample_p_row<DT>(prow, ro[i], V, c, greedy, u, sc);
      if (threadIdx.x == 0) {
        const int tok = (int)min(max(toks[i], 0ll), (long long)V - 1);
        s_acc[i] = (s == toks[i]);
        job.p_tok[(long long)b * g + i] = row_prob<DT>(ro[i], prow, tok, c);
        job.q_tok[(long long)b * g + i] = 0.0f;
      }
    }
  } else if (threadIdx.x < g) {
    const int i = threadIdx.x;
    const int tok = (int)min(max(toks[i], 0ll), (long long)V - 1);  // ids outside [0,V) are clamped
    const void* prow = row_ptr<DT>(rj, (long long)b * rps + i);
    const void* qrow = row_ptr<DT>(rj, (long long)b * rps + rj.nT + i);
    const float p = row_prob<DT>(ro[i], prow, tok, c);
    const float q = row_prob<DT>(ro[rj.nT + i], qrow, tok, c);
    const float u = job.u_accept ? job.u_accept[(long long)b * g + i] : philox_uniform(job.seed, job_offset(job), seq, i);
    int acc;
    if (job.flags & SPECDEC_ACCEPT_BATCHED) {  // engine/infer_engine.py:303-305 (python floats = double)
      const double ap = (q <= 0.0f) ? 1.0 : fmin(1.0, (double)p / (double)q);
      acc = ((double)u < ap);
    } else {  // sampling/speculative_decoding.py:143 : reject iff r > p/q  (NaN => accept)
      const float frac = __fdiv_rn(p, q);
      acc = !(u > frac);
    }
    s_acc[i] = acc;
    job.p_tok[(long long)b * g + i] = p;
    job.q_tok[(long long)b * g + i] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = g;
    for (int i = 0; i < g; ++i) {
      job.mask[(long long)b * g + i] = (unsigned char)s_acc[i];
      if (!s_acc[i] && n == g) n = i;
    }
    const int fs = first_stop_index(toks, n, job.stop, job.n_stop, job.flags);
    job.n_acc[b] = n;
    job.first_stop[b] = fs;
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;

  // 2. next token
  const float us = job.u_sample ? job.u_sample[b] : philox_uniform(job.seed, job_offset(job), seq, (unsigned)job.lane_sample);
  long long x = -1;
  int prow_idx = -1;  // target row x was drawn from (for next_prob)
  if (n == g) {
    if (!(job.flags & SPECDEC_NO_BONUS)) {
      prow_idx = g;
      x = sample_p_row<DT>(row_ptr<DT>(rj, (long long)b * rps + g), ro[g], V, c, greedy, us, sc);
    }
  } else {
    const void* prow = row_ptr<DT>(rj, (long long)b * rps + n);
    if (ngram || (job.flags & SPECDEC_SKIP_ADJUST)) {
      prow_idx = n;
      x = sample_p_row<DT>(prow, ro[n], V, c, greedy, us, sc);
    } else {
      const void* qrow = row_ptr<DT>(rj, (long long)b * rps + rj.nT + n);
      const u64 rmin = (job.flags & SPECDEC_RESID_FALLBACK) ? 1152921ull : 0ull;
      x = sample_residual<DT>(prow, ro[n], qrow, ro[rj.nT + n], V, c, greedy, us, rmin, sc);
      if (x < 0) {
        prow_idx = n;
        x = sample_p_row<DT>(prow, ro[n], V, c, greedy, us, sc);
      }
    }
  }
  if (threadIdx.x == 0) {
    job.next_tok[b] = x;
    if (job.next_prob)
      job.next_prob[b] = (prow_idx >= 0 && x >= 0)
                             ? row_prob<DT>(ro[prow_idx], row_ptr<DT>(rj, (long long)b * rps + prow_idx), (int)x, c)
                             : 0.0f;
    if (job.packed) {
      int* pk = job.packed + (long long)b * (g + 2);
      pk[0] = n;
      for (int i = 0; i < g + 1; ++i) pk[1 + i] = -1;
      for (int i = 0; i < n; ++i) pk[1 + i] = (int)toks[i];
      pk[1 + n] = (int)x;
    }
  }
}

#include "hybrid.cuh"
#include "cluster_small.cuh"
#include "rowsel_tma.cuh"

// ---------------------------------------------------------------------------------------------
// probabilities materialised (LogitsProcessor.__call__), sample() on given probs, philox dump
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(NT, 1) probs_kernel(RowJob job, float* probs, float* stats) {
  for (long long r = blockIdx.x; r < job.R; r += gridDim.x) {
    const void* row = row_ptr<DT>(job, r);
    const bool aligned = (((size_t)row) & 15) == 0;
    const RowOut ro = job.out[r];
    float* out = probs ? probs + (size_t)r * job.V : nullptr;
    if (out)
      sweep<DT>(row, job.V, aligned, [&](const float(&x)[8], int j0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j0 + k < job.V)
            out[j0 + k] = __fmul_rn(kept(ro, x[k], j0 + k) ? cweight(x[k], job.c, ro.mc) : 0.0f, ro.inv);
      });
    if (stats && threadIdx.x == 0) {
      float* s = stats + r * 8;
      s[0] = ro.m; s[1] = __fmul_rn(__ull2float_rn(ro.Sfix), 0x1p-40f); s[2] = ro.inv; s[3] = ro.cut;
      s[4] = (float)ro.jcut; s[5] = 0; s[6] = 0; s[7] = 0;
    }
  }
}

__global__ void __launch_bounds__(NT, 1) sample_probs_kernel(const float* probs, long long rows, int V, int greedy,
                                                             const float* u, long long* tok) {
  __shared__ u64 part[MAXPART];
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ int shi[33];
  __shared__ long long s_res;
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* row = probs + (size_t)r * V;
    const bool aligned = (((size_t)row) & 15) == 0;
    long long x;
    if (greedy) {
      float best = -INFINITY;
      int idx = 0x7FFFFFFF;
      sweep<DT_F32>(row, V, aligned, [&](const float(&p)[8], int j0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j0 + k < V && p[k] > best) { best = p[k]; idx = j0 + k; }
      });
      x = block_argmax(best, idx, shf, shi, &s_res);
      if (x < 0) x = 0;
    } else {
      u64 total;
      auto wf = [&](int v, u64(&w)[8]) {
        float p[8];
        load8<DT_F32>(row, v, V, aligned, p);
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = (v * 8 + k < V && p[k] > 0.0f) ? fix40(p[k]) : 0ull;
      };
      x = invcdf_sweep(V, wf, false, 0ull, u24_of(u[r]), part, sh64, &s_res, total);
      if (x < 0) x = 0;
    }
    if (threadIdx.x == 0) tok[r] = x;
  }
}

__global__ void philox_kernel(u64 seed, u64 offset, long long seq0, int B, int g, float* ua, float* us) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * (g + 1)) return;
  const int b = t / (g + 1), i = t % (g + 1);
  const unsigned seq = (unsigned)(seq0 + b);
  if (i < g) { if (ua) ua[(long long)b * g + i] = philox_uniform(seed, offset, seq, (unsigned)i); }
  else if (us) us[b] = philox_uniform(seed, offset, seq, 0x10000u);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Process-wide state of the library: the tuning / test options below, the optional profiling events and the per-device
// caches (function attributes, the auxiliary stream and its fork/join events).  Every entry point that launches work
// or changes that state holds g_api_mu for the duration of its (host-side, microseconds) enqueue, so concurrent
// calls from several host threads are serialised at the API boundary instead of racing on it.
static std::mutex g_api_mu;
static cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};  // optional: start / after rowstats / after decide
static int g_no_fast_ngram = 0;    // test hook: specdec_set_option("no_fast_ngram", 1)
static int g_no_fast_nucleus = 0;  // test hook: specdec_set_option("no_fast_nucleus", 1)
static int g_no_tma_nucleus = 1;   // specdec_set_option("no_tma_nucleus", 0) => first sweep of the top-p rows as a TMA row-kernel launch
                                   // (off: measured 1.42 -> 1.38 ms on flat rows but 0.55 -> 0.60 ms on LLM-like rows, whose
                                   // second sweep then comes from HBM instead of L2)
static int g_no_hist_nucleus = 0;  // test hook: specdec_set_option("no_hist_nucleus", 1) => band search for flat rows
static int g_force_ldg = 0;  // test hook: specdec_set_option("force_ldg", 1)
// programmatic dependent launch is used for eager launches only: inside a stream capture the programmatic edges made
// the replayed graph slower (0.225 vs 0.189 ms per step), so captured calls launch in plain stream order
static bool pdl_enabled(cudaStream_t st);
static int g_no_pdl = 0;         // test hook: specdec_set_option("no_pdl", 1) => plain stream-ordered launches of plan / tail
static int g_tma_ngram = 1;      // greedy n-gram verify on 16-bit rows: arg-max from the TMA row pipeline ("tma_ngram"=0: LDG kernel)
static bool pdl_enabled(cudaStream_t st) {
  if (g_no_pdl) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone;
}
static int g_split_lists = 0;    // masked modes: plan and the kept-token-list draw as two launches (test hook)
static int g_static_rows = 0;    // row kernel: rows assigned by blockIdx instead of claimed from a counter (test hook)
static int g_tf_balance = 1;     // fused tail: slice sizes rounded to a multiple of 8 segments (one per warp)
static int g_tail_slots = 1;     // fused tail: inter-CTA exchange through self-validating words (0: atomics + counters)
static int g_no_fused_tail = 0;  // test hook: specdec_set_option("no_fused_tail", 1) => exact_rows + sample_partial
static int g_no_klist = 0;       // test hook: specdec_set_option("no_klist", 1) => masked modes always draw by a sweep over the row
// per-device caches (function attributes are per device; one process may drive several GPUs)
constexpr int MAXDEV = 32;
static int cur_dev() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < MAXDEV) ? dev : 0;
}
static int g_sms = 0;
static int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

static int fill_rowjob(RowJob& rj, const void* tgt, const void* drf, long long tsb, long long tsg, long long dsb,
                       long long dsg, int nT, int nD, int V, float temperature, int top_k, float top_p, long long R,
                       void* workspace, size_t workspace_bytes) {
  if (V <= 0 || V > MAXPART * 256 || !(temperature > 0.0f) || R < 0) return SPECDEC_ERR_RANGE;
  if (workspace_bytes < (size_t)R * sizeof(RowOut) || (R > 0 && !workspace)) return SPECDEC_ERR_WORKSPACE;
  rj.tgt = tgt; rj.drf = drf; rj.tsb = tsb; rj.tsg = tsg; rj.dsb = dsb; rj.dsg = dsg;
  rj.nT = nT; rj.nD = nD; rj.V = V;
  rj.c = (float)(1.4426950408889634 / (double)temperature);
  rj.c1 = (float)1.4426950408889634;
  rj.top_k = (top_k > 0 && top_k < V) ? top_k : 0;
  rj.use_p = (top_p > 0.0f && top_p < 1.0f) ? 1 : 0;
  rj.tpq = rj.use_p ? (u64)((double)top_p * 4294967296.0) : 0ull;
  rj.R = R;
  rj.out = (RowOut*)workspace;
  rj.skip_resolved = 0;
  rj.pre_stats = 0;
  rj.klist = nullptr;
  rj.n_unres = nullptr;
  return 0;
}

// Top-p rows: the first sweep of nucleus_fast_kernel (row max, MUFU T=1 mass, candidate threshold) as a launch of the
// TMA row pipeline (hybrid.cuh / rowfast_tma.cuh).  Returns false when the rows are not TMA-eligible.
template <int DT>
static bool nucleus_prepass_tma(const RowJob& rj, cudaStream_t st);

static int g_rowsel_probe = 0;  // timing probe: the selector warp skips the selection (rows fall through)
static int g_no_rowsel = 0;  // test hook: specdec_set_option("no_rowsel", 1) => masked modes without the streamed selection kernel

// masked modes on TMA-eligible 16-bit rows: rowsel_tma_kernel resolves the rows at HBM speed; false = not eligible
template <int DT, bool HK, bool HP>
static bool launch_rowsel(const RowJob& rj, cudaStream_t st) {
  if (DT == DT_F32) return false;
  const size_t es = 2;
  const bool ok = !g_no_rowsel && !g_no_fast_nucleus && !g_force_ldg && rj.R > 0 && (!HK || rj.top_k <= RS2_CONSUMERS) &&
                  (((size_t)rj.tgt | (size_t)rj.drf) & 15) == 0 && ((size_t)rj.V * es) % 16 == 0 &&
                  ((size_t)rj.tsb * es) % 16 == 0 && ((size_t)rj.tsg * es) % 16 == 0 && ((size_t)rj.dsb * es) % 16 == 0 &&
                  ((size_t)rj.dsg * es) % 16 == 0;
  if (!ok) return false;
  static int occ_dev[MAXDEV];
  int& occ = occ_dev[cur_dev()];
  if (!occ) {
    if (cudaFuncSetAttribute(rowsel_tma_kernel<DT, HK, HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS2_SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rowsel_tma_kernel<DT, HK, HP>, RS2_THREADS, RS2_SMEM) != cudaSuccess || occ < 1)
      occ = -1;
  }
  if (occ < 1) return false;
  const long long cap = (long long)(occ < 2 ? occ : 2) * num_sms();
  RowJob rjp = rj;
  rjp.pre_stats = g_rowsel_probe;
  rowsel_tma_kernel<DT, HK, HP><<<(unsigned)(rj.R < cap ? rj.R : cap), RS2_THREADS, RS2_SMEM, st>>>(rjp);
  return cudaGetLastError() == cudaSuccess;
}

template <int DT>
static cudaError_t launch_rowstats(const RowJob& rj, cudaStream_t st) {
  if (rj.R == 0) return cudaSuccess;
  const bool masked = rj.top_k > 0 || rj.use_p;
  const size_t smem = masked ? CAND_SMEM : 0;
  const long long cap = 2LL * num_sms();
  const int grid = (int)(rj.R < cap ? rj.R : cap);
#define RS_LAUNCH(HKv, HPv, JOB)                                                                                    \
  do {                                                                                                              \
    if (masked) {                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(rowstats_kernel<DT, HKv, HPv>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem);                                                              \
      if (e != cudaSuccess) return e;                                                                               \
    }                                                                                                               \
    rowstats_kernel<DT, HKv, HPv><<<grid, RS_NT, smem, st>>>(JOB);                                                   \
  } while (0)
  if (!masked) { RS_LAUNCH(false, false, rj); return cudaGetLastError(); }
  // Fast route: rowsel_tma_kernel (one streamed pass + gather of the hot slices + exact selection by a selector warp);
  // the rows it leaves unresolved go to the exact kernels below, which skip every row already flagged.
  // (n_unres counts the rows rowsel_tma_kernel left: it only means something when that kernel ran, and for pure top-p
  // the rows pass through two more resolvers before the fallback -- there the fallback polls the flags as before)
  RowJob left = rj;
  if (rj.top_k > 0 && rj.use_p) {
    if (launch_rowsel<DT, true, true>(rj, st)) left.skip_resolved = 1;
    else left.n_unres = nullptr;
    RS_LAUNCH(true, true, left);
  } else if (rj.top_k > 0) {
    if (launch_rowsel<DT, true, false>(rj, st)) left.skip_resolved = 1;
    else left.n_unres = nullptr;
    RS_LAUNCH(true, false, left);
  } else {
    left.n_unres = nullptr;
    if (!g_no_fast_nucleus) {
      cudaError_t e;
      if (!launch_rowsel<DT, false, true>(rj, st)) {
        if ((e = cudaFuncSetAttribute(nucleus_fast_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        RowJob rj1 = rj;
        if (nucleus_prepass_tma<DT>(rj, st)) rj1.pre_stats = 1;  // max / MUFU mass / candidate threshold at HBM speed
        nucleus_fast_kernel<DT><<<grid, RS_NT, smem, st>>>(rj1);
      }
      if (!g_no_hist_nucleus) {  // flat rows: radix-select by private histograms, cost independent of the nucleus size
        if ((e = cudaFuncSetAttribute(nucleus_hist_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NH_SMEM)) != cudaSuccess) return e;
        const int gh = (int)(rj.R < (long long)num_sms() ? rj.R : (long long)num_sms());
        nucleus_hist_kernel<DT><<<gh, NH_T, NH_SMEM, st>>>(rj);
      }
      left.skip_resolved = 1;
    }
    RS_LAUNCH(false, true, left);
  }
#undef RS_LAUNCH
  return cudaGetLastError();
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr int CS_MAXCL = 16;  // largest cluster of the cluster-per-sequence path (cluster_small.cuh)
struct WsLayout {
  size_t rowout, zero, zero_bytes, zero_bytes_mega, tasks, status, samp, part, rpart, xs, dbg, klist, cpart, total;
  int nseg_pad, xs_stride;
};
static WsLayout ws_layout(long long B, int gamma, int V, long long R) {
  WsLayout w;
  const int NV = (V + 7) >> 3;
  w.nseg_pad = ((NV + 31) >> 5) + 1;
  w.xs_stride = (w.nseg_pad - 1 + MG_MIN_SPC - 1) / MG_MIN_SPC;  // exact items per sequence at the smallest item size
  if (w.xs_stride < 64) w.xs_stride = 64;                         // (tail_slots_kernel: up to 64 slices per sequence)
  size_t o = 0;
  w.rowout = o; o = al256(o + (size_t)R * sizeof(RowOut));
  w.zero = o;  // ---- zeroed at the head of every call
  o += (size_t)R * 8 + (size_t)B * 8 * 2 + (size_t)B * 8 * 2 + 64 + 64 + (size_t)B * 4 * 7 + (size_t)MG_SM_SLOTS * 4;
  w.zero_bytes = o - w.zero;
  o = al256(o);
  // ---- zeroed as well when the megakernel runs (its CTAs exchange self-validating words, mega.cuh)
  w.rpart = o; o = al256(o + (size_t)R * MG_MAXU * sizeof(float2));  // per row slice (max, sum)
  w.xs = o; o = al256(o + (size_t)B * w.xs_stride * 32);              // per exact item {Sp, Sq, greedy key, -}
  w.part = o; o = al256(o + (size_t)B * w.nseg_pad * 8);
  w.zero_bytes_mega = o - w.zero;
  w.tasks = o; o = al256(o + (size_t)B * (gamma > 0 ? gamma : 1) * 4);
  w.status = o; o = al256(o + (size_t)B * (gamma > 0 ? gamma : 1));
  w.samp = o; o = al256(o + (size_t)B * 4 * SAMP_N);
  w.dbg = o; o = al256(o + (size_t)(16 + B * 8 + 1024) * 8);          // megakernel: debug timeline + SM of every CTA
  w.klist = o; o = al256(o + (size_t)R * KL_MAX * sizeof(int2));      // masked modes: kept-token lists of small kept sets
  w.cpart = o; o = al256(o + (size_t)R * CS_MAXCL * sizeof(float2));  // cluster-per-sequence path: (max, sum) of every row slice
  w.total = o;
  return w;
}
static HybridWs ws_pointers(const WsLayout& w, void* workspace, long long B, long long R) {
  char* base = (char*)workspace;
  HybridWs h;
  h.acc = (u64*)(base + w.zero);
  h.tot = h.acc + R;
  h.best = h.tot + B;
  h.acc2 = h.best + B;
  h.ntasks = (int*)(h.acc2 + 2 * B);
  h.ticket = h.ntasks + 16;
  h.rows_done = h.ticket + 16;
  h.seq_tasks = h.rows_done + B;
  h.exact_done = h.seq_tasks + B;
  h.part_done = h.exact_done + B;
  h.fin_done = h.part_done + B;
  h.decided = h.fin_done + B;
  h.plan_done = h.decided + B;
  h.sm_slots = h.plan_done + B;
  h.r_ticket = h.ntasks + 11;
  h.abort = h.ntasks + 10;   // (ntasks[0..7] / ticket[0..7]: one counter per chunk)
  h.r_claim = g_static_rows ? nullptr : h.ticket + 8;  // ticket[8..15]: row-claim counter per chunk
  h.x_next = h.ntasks + 8;
  h.p_next = h.ntasks + 9;
  h.rpart = (float2*)(base + w.rpart);
  h.xs = (u64*)(base + w.xs);
  h.xs_stride = w.xs_stride;
  h.dbg = nullptr;
  h.fused = 0;
  h.tasks = (int*)(base + w.tasks);
  h.status = (unsigned char*)(base + w.status);
  h.samp = (int*)(base + w.samp);
  h.part = (u64*)(base + w.part);
  h.nseg_pad = w.nseg_pad;
  return h;
}

static int g_chunks = 2;    // batch chunks pipelined on two streams (specdec_set_option("chunks", n); 1 = off)
static int g_chunk0_pct = 50;  // share of the batch in chunk 0 when chunks == 2
static int g_p1_ctas = 3;   // row-kernel CTAs per SM while a tail kernel of the previous chunk shares the SMs
static int g_tf_ch = TF_CH_DEFAULT;  // CTAs per sequence of tail_fused_kernel
static int g_mega = 0;         // "mega"=1: plain modes on 16-bit TMA-eligible rows as ONE persistent cooperative launch (mega.cuh).
                               // Off by default: measured 0.26 ms vs 0.19 ms for the three-launch pipeline at the headline shape
                               // (139 M vs 114 M warp instructions, the exact items are latency-bound) -- see DESIGN.md 4.3
static int g_mega_r = 2;       // R (row streaming) CTAs per SM of the megakernel; the others are X (exact) CTAs
static int g_mega_unit = 4;    // TMA stages (16 KB) per row slice
static int g_mega_spc = 24;    // 256-element segments per exact item (2 KB of cached weights each, <= 24)
static int g_mega_keep_l2 = 1; // rows streamed with the normal L2 policy (the deciding pair is re-read from L2)
static int g_mega_dbg = 0;     // record the %globaltimer timeline of the megakernel (specdec_debug_timeline)
struct AuxStream {  // per device: the library's high-priority stream for the chunk pipeline + fork/join events
  cudaStream_t stream;
  cudaEvent_t ev_a[8], ev_b;
};
static AuxStream g_aux[MAXDEV];

template <int DT>
static bool nucleus_prepass_tma(const RowJob& rj, cudaStream_t st) {
  const size_t es = (DT == DT_F32) ? 4 : 2;
  const bool ok = DT != DT_F32 && !g_no_tma_nucleus && !g_force_ldg && rj.R > 0 && (((size_t)rj.tgt | (size_t)rj.drf) & 15) == 0 &&
                  ((size_t)rj.V * es) % 16 == 0 && ((size_t)rj.tsb * es) % 16 == 0 && ((size_t)rj.tsg * es) % 16 == 0 &&
                  ((size_t)rj.dsb * es) % 16 == 0 && ((size_t)rj.dsg * es) % 16 == 0;
  if (!ok) return false;
  static int occ_dev[MAXDEV];
  int& occ = occ_dev[cur_dev()];
  if (!occ) {
    if (cudaFuncSetAttribute(rowfast_tma_kernel<DT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM) != cudaSuccess) return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rowfast_tma_kernel<DT, 1>, TS_THREADS, TS_SMEM) != cudaSuccess || occ < 1)
      occ = 1;
  }
  DecideJob dj;
  memset(&dj, 0, sizeof(dj));
  dj.rj = rj;
  dj.rj.c = rj.c1;  // T = 1 masses
  HybridWs ws;
  memset(&ws, 0, sizeof(ws));
  const long long cap = (long long)(occ < 4 ? occ : 4) * num_sms();
  rowfast_tma_kernel<DT, 1><<<(unsigned)(rj.R < cap ? rj.R : cap), TS_THREADS, TS_SMEM, st>>>(dj, ws);
  return cudaGetLastError() == cudaSuccess;
}

// greedy n-gram verify: max / MUFU sum / first index of the maximum of every target row through the TMA row pipeline
template <int DT>
static bool ngram_argmax_tma(const RowJob& rj, cudaStream_t st) {
  if (DT == DT_F32 || !g_tma_ngram) return false;
  const size_t es = 2;
  const bool ok = !g_force_ldg && rj.R > 0 && (((size_t)rj.tgt) & 15) == 0 && ((size_t)rj.V * es) % 16 == 0 &&
                  ((size_t)rj.tsb * es) % 16 == 0 && ((size_t)rj.tsg * es) % 16 == 0 && rj.nD == 0;
  if (!ok) return false;
  static int occ_dev[MAXDEV];
  int& occ = occ_dev[cur_dev()];
  if (!occ) {
    if (cudaFuncSetAttribute(rowfast_tma_kernel<DT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM) != cudaSuccess) return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rowfast_tma_kernel<DT, 2>, TS_THREADS, TS_SMEM) != cudaSuccess || occ < 1)
      occ = 1;
  }
  DecideJob dj;
  memset(&dj, 0, sizeof(dj));
  dj.rj = rj;
  HybridWs ws;
  memset(&ws, 0, sizeof(ws));
  const long long cap = (long long)(occ < 4 ? occ : 4) * num_sms();
  rowfast_tma_kernel<DT, 2><<<(unsigned)(rj.R < cap ? rj.R : cap), TS_THREADS, TS_SMEM, st>>>(dj, ws);
  return cudaGetLastError() == cudaSuccess;
}

// phase A: row statistics of the job's rows.  limit_ctas > 0 caps the persistent grid per SM.
template <int DT>
static cudaError_t launch_phase_a(const DecideJob& dj, const HybridWs& ws_in, int B, int ctas_per_sm, cudaStream_t st,
                                  bool overlap_prev = false) {
  const RowJob& rj = dj.rj;
  HybridWs ws = ws_in;
  const bool masked = rj.top_k > 0 || rj.use_p;
  cudaError_t e;
  if (masked) return launch_rowstats<DT>(rj, st);
  if (dj.gamma == 0) {
    rowfast_kernel<DT><<<(unsigned)rj.R, FT, 0, st>>>(dj, ws);
    return cudaGetLastError();
  }
  const size_t es = (DT == DT_F32) ? 4 : 2;
  const bool tma_ok = DT != DT_F32 && (((size_t)rj.tgt | (size_t)rj.drf) & 15) == 0 && ((size_t)rj.V * es) % 16 == 0 &&
                      ((size_t)rj.tsb * es) % 16 == 0 && ((size_t)rj.tsg * es) % 16 == 0 &&
                      ((size_t)rj.dsb * es) % 16 == 0 && ((size_t)rj.dsg * es) % 16 == 0 && !g_force_ldg;
  if (tma_ok) {
    static int occ_dev[MAXDEV];  // resident CTAs per SM of the persistent grid (0 = attribute not set on this device yet)
    int& occ = occ_dev[cur_dev()];
    if (!occ) {
      e = cudaFuncSetAttribute(rowfast_tma_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
      if (e != cudaSuccess) return e;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rowfast_tma_kernel<DT>, TS_THREADS, TS_SMEM) != cudaSuccess || occ < 1)
        occ = 1;
    }
    const int per_sm = (ctas_per_sm + 1) < occ ? (ctas_per_sm + 1) : occ;
    const long long cap = (long long)per_sm * num_sms();
    // rows are claimed from a counter only when a CTA gets more than one (otherwise the claim's round trip is pure
    // latency: 93 -> 106 us at B = 128, one row per CTA)
    if (rj.R <= cap) ws.r_claim = nullptr;
    if (overlap_prev && pdl_enabled(st)) {
      // The row kernel of chunk i > 0 depends on nothing the row kernel of chunk i-1 does (other rows, other
      // RowOut records): launched with programmatic stream serialization and NO dependency wait, its CTAs start
      // as the previous row kernel's CTAs drain instead of after its last CTA.
      cudaLaunchAttribute pdl[1];
      pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      pdl[0].val.programmaticStreamSerializationAllowed = 1;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.stream = st; cfg.attrs = pdl; cfg.numAttrs = 1;
      cfg.gridDim = dim3((unsigned)(rj.R < cap ? rj.R : cap)); cfg.blockDim = dim3(TS_THREADS); cfg.dynamicSmemBytes = TS_SMEM;
      return cudaLaunchKernelEx(&cfg, rowfast_tma_kernel<DT>, dj, ws);
    }
    rowfast_tma_kernel<DT><<<(unsigned)(rj.R < cap ? rj.R : cap), TS_THREADS, TS_SMEM, st>>>(dj, ws);
  } else {
    rowfast_kernel<DT><<<(unsigned)rj.R, FT, 0, st>>>(dj, ws);
  }
  return cudaGetLastError();
}

// phase B: plan, exact sums of the deciding rows, sampling sweep + finalize
template <int DT>
static cudaError_t launch_phase_b(const DecideJob& dj, const HybridWs& ws_in, int B, cudaStream_t st) {
  const RowJob& rj = dj.rj;
  const bool masked = rj.top_k > 0 || rj.use_p;
  HybridWs ws = ws_in;
  // fused tail (tail_fused.cuh): nch CTAs per sequence keep the canonical weights of the deciding row pair
  // in shared memory; needs the slice of one CTA to fit
  const int nseg = ((((rj.V + 7) >> 3) + 31) >> 5);
  int nch = g_tf_ch, spc = (nseg + nch - 1) / nch;
  if (g_tf_balance && spc >= 8) {
    // slices of a multiple of 8 segments: the 8 warps of a CTA get the same number of 256-pair segments (26 segments
    // per CTA left 6 of the 8 warps idle for a quarter of both phases)
    spc = ((spc + 4) / 8) * 8;
    nch = (nseg + spc - 1) / spc;
  }
  const size_t tf_smem = (size_t)spc * TF_SEG_BYTES;
  bool fused_ok = !masked && dj.gamma > 0 && tf_smem <= 200 * 1024 && !g_no_fused_tail;
  // exchange between the CTAs of a sequence: self-validating words (tail_slots_kernel, default) or atomics + counters
  const bool slots = g_tail_slots && nch <= ws.xs_stride;
  auto tail = slots ? (dj.greedy ? tail_slots_kernel<DT, true> : tail_slots_kernel<DT, false>)
                    : (dj.greedy ? tail_fused_kernel<DT, true> : tail_fused_kernel<DT, false>);
  if (fused_ok) {
    // The nch CTAs of a sequence wait for one another (tail_fused.cuh): that needs at least nch CTAs co-resident on
    // the device (ticket order then guarantees progress).  Under an SM limit (MPS / green contexts) or with a huge
    // slice the occupancy query says otherwise and the step takes the split pipeline, whose CTAs never wait.
    static size_t attr_smem_dev[MAXDEV][4];  // dynamic shared memory already granted (per device, per instantiation)
    static size_t occ_smem_dev[MAXDEV][4];
    static int occ_dev[MAXDEV][4];
    const int dv = cur_dev(), gi = (dj.greedy ? 1 : 0) + (slots ? 2 : 0);
    if (attr_smem_dev[dv][gi] < tf_smem) {
      cudaError_t e = cudaFuncSetAttribute(tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf_smem);
      if (e != cudaSuccess) return e;
      attr_smem_dev[dv][gi] = tf_smem;
    }
    if (occ_smem_dev[dv][gi] != tf_smem) {
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tail, TF_T, tf_smem) != cudaSuccess) occ = 0;
      occ_dev[dv][gi] = occ;
      occ_smem_dev[dv][gi] = tf_smem;
    }
    if ((long long)occ_dev[dv][gi] * num_sms() < nch) fused_ok = false;
  }
  if (fused_ok) {
    ws.fused = 1;
    cudaError_t e;
    // programmatic dependent launches: the CTAs of plan / tail are scheduled while their predecessor drains
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = pdl_enabled(st) ? 1 : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.stream = st; cfg.attrs = pdl; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)((B + 7) / 8)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0;
    if ((e = cudaLaunchKernelEx(&cfg, plan_kernel<DT>, dj, ws)) != cudaSuccess) return e;
    cfg.gridDim = dim3((unsigned)B * nch); cfg.blockDim = dim3(TF_T); cfg.dynamicSmemBytes = tf_smem;
    if ((e = cudaLaunchKernelEx(&cfg, tail, dj, ws, spc, nch)) != cudaSuccess) return e;
    return cudaGetLastError();
  }
  // split pipeline (masked modes, gamma == 0, test hook): plan, [exact sums of the task rows], sampling sweep;
  // programmatic dependent launches like the fused path
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl[0].val.programmaticStreamSerializationAllowed = pdl_enabled(st) ? 1 : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.stream = st; cfg.attrs = pdl; cfg.numAttrs = 1; cfg.dynamicSmemBytes = 0;
  cudaError_t e;
  cfg.gridDim = dim3((unsigned)((B + 7) / 8)); cfg.blockDim = dim3(256);
  // masked modes with kept-token lists: the plan decides every sequence itself (no exact tasks), so the same warp draws
  // from the lists right away (plan_lists_kernel); "split_lists" = 1 keeps the two launches (test hook)
  const bool lists = masked && rj.klist != nullptr;
  if (lists && !g_split_lists) {
    if ((e = cudaLaunchKernelEx(&cfg, plan_lists_kernel<DT>, dj, ws)) != cudaSuccess) return e;
  } else {
    if ((e = cudaLaunchKernelEx(&cfg, plan_kernel<DT>, dj, ws)) != cudaSuccess) return e;
  }
  cfg.gridDim = dim3((unsigned)B, CH); cfg.blockDim = dim3(PT);
  if (!masked && dj.gamma > 0)  // tasks are looped over
    if ((e = cudaLaunchKernelEx(&cfg, exact_rows_kernel<DT>, dj, ws)) != cudaSuccess) return e;
  if (lists && g_split_lists) {  // sequences whose deciding rows carry kept-token lists are drawn without a sweep
    cfg.gridDim = dim3((unsigned)((B + SL_WARPS - 1) / SL_WARPS)); cfg.blockDim = dim3(SL_WARPS * 32);
    if ((e = cudaLaunchKernelEx(&cfg, sample_lists_kernel<DT>, dj, ws, B)) != cudaSuccess) return e;
    cfg.gridDim = dim3((unsigned)B, CH); cfg.blockDim = dim3(PT);
  }
  auto samp = masked ? (dj.greedy ? sample_partial_kernel<DT, true, true> : sample_partial_kernel<DT, true, false>)
                     : (dj.greedy ? sample_partial_kernel<DT, false, true> : sample_partial_kernel<DT, false, false>);
  if ((e = cudaLaunchKernelEx(&cfg, samp, dj, ws)) != cudaSuccess) return e;
  return cudaGetLastError();
}

// the sub-job of sequences [b0, b0 + nb)
template <int DT>
static void sub_job(const DecideJob& dj, const HybridWs& ws, int b0, int nb, int half, DecideJob& o, HybridWs& w) {
  o = dj; w = ws;
  const size_t es = (DT == DT_F32) ? 4 : 2;
  const int g = dj.gamma, rps = dj.rj.nT + dj.rj.nD;
  o.rj.tgt = (const char*)dj.rj.tgt + (size_t)b0 * dj.rj.tsb * es;
  if (dj.rj.drf) o.rj.drf = (const char*)dj.rj.drf + (size_t)b0 * dj.rj.dsb * es;
  o.rj.R = (long long)nb * rps;
  o.rj.out = dj.rj.out + (size_t)b0 * rps;
  if (dj.draft_tokens) o.draft_tokens = dj.draft_tokens + (size_t)b0 * g;
  if (dj.u_accept) o.u_accept = dj.u_accept + (size_t)b0 * g;
  if (dj.u_sample) o.u_sample = dj.u_sample + b0;
  o.seq0 = dj.seq0 + b0;
  o.n_acc = dj.n_acc + b0; o.next_tok = dj.next_tok + b0; o.first_stop = dj.first_stop + b0;
  if (dj.mask) o.mask = dj.mask + (size_t)b0 * g;
  if (dj.p_tok) o.p_tok = dj.p_tok + (size_t)b0 * g;
  if (dj.q_tok) o.q_tok = dj.q_tok + (size_t)b0 * g;
  if (dj.next_prob) o.next_prob = dj.next_prob + b0;
  if (dj.packed) o.packed = dj.packed + (size_t)b0 * (g + 2);
  w.acc = ws.acc + (size_t)b0 * rps; w.tot = ws.tot + b0; w.best = ws.best + b0;
  w.acc2 = ws.acc2 + (size_t)b0 * 2; w.fin_done = ws.fin_done + b0; w.decided = ws.decided + b0;
  w.ticket = ws.ticket + half;
  if (ws.r_claim) w.r_claim = ws.r_claim + half;
  w.ntasks = ws.ntasks + half;
  w.tasks = ws.tasks + (size_t)b0 * (g > 0 ? g : 1);
  w.status = ws.status + (size_t)b0 * (g > 0 ? g : 1);
  w.samp = ws.samp + (size_t)b0 * SAMP_N;
  w.rows_done = ws.rows_done + b0; w.seq_tasks = ws.seq_tasks + b0;
  w.exact_done = ws.exact_done + b0; w.part_done = ws.part_done + b0;
  w.part = ws.part + (size_t)b0 * ws.nseg_pad;
  w.xs = ws.xs + (size_t)b0 * ws.xs_stride * 4;
}

// The plain modes on TMA-eligible 16-bit rows: one persistent cooperative launch (mega.cuh).  mega_config() returns
// false when the shape is not eligible (the caller then takes the three-launch pipeline).
template <int DT>
static bool mega_config(const DecideJob& dj, int B, MegaCfg& cfg, int& grid) {
  const RowJob& rj = dj.rj;
  const size_t es = (DT == DT_F32) ? 4 : 2;
  const bool ok = g_mega && DT != DT_F32 && dj.gamma > 0 && rj.top_k == 0 && !rj.use_p && !g_force_ldg && !g_no_fused_tail &&
                  !(dj.flags & SPECDEC_NGRAM) && (((size_t)rj.tgt | (size_t)rj.drf) & 15) == 0 && ((size_t)rj.V * es) % 16 == 0 &&
                  ((size_t)rj.tsb * es) % 16 == 0 && ((size_t)rj.tsg * es) % 16 == 0 && ((size_t)rj.dsb * es) % 16 == 0 &&
                  ((size_t)rj.dsg * es) % 16 == 0;
  if (!ok) return false;
  static int occ_dev[MAXDEV][2];
  int& occ = occ_dev[cur_dev()][dj.greedy ? 1 : 0];
  if (!occ) {
    auto kern = dj.greedy ? verify_mega_kernel<DT, true> : verify_mega_kernel<DT, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TS_THREADS, TS_SMEM) != cudaSuccess || occ < 1)
      occ = -1;
  }
  if (occ < 2) return false;
  const int per_sm = occ < 4 ? occ : 4;
  const int r_per_sm = g_mega_r < per_sm ? g_mega_r : per_sm - 1;
  grid = per_sm * num_sms();
  cfg.n_r = r_per_sm * num_sms();
  cfg.n_x = grid - cfg.n_r;
  cfg.r_per_sm = r_per_sm;
  const long long row_bytes = (long long)rj.V * (long long)es;
  const int nst = (int)((row_bytes + TS_STAGE_BYTES - 1) / TS_STAGE_BYTES);
  cfg.unit_stages = g_mega_unit;
  if ((nst + cfg.unit_stages - 1) / cfg.unit_stages > MG_MAXU) cfg.unit_stages = (nst + MG_MAXU - 1) / MG_MAXU;
  cfg.U = (nst + cfg.unit_stages - 1) / cfg.unit_stages;
  const int nseg = ((((rj.V + 7) >> 3) + 31) >> 5);
  cfg.spc = nseg < g_mega_spc ? nseg : g_mega_spc;
  cfg.S = (nseg + cfg.spc - 1) / cfg.spc;
  cfg.B = B;
  cfg.keep_l2 = g_mega_keep_l2;
  return grid - cfg.n_r >= 3 * cfg.S + 1;  // forward-progress condition of the item claims
}
template <int DT>
static cudaError_t launch_mega(const DecideJob& dj, HybridWs ws, const MegaCfg& cfg, int grid, cudaStream_t st) {
  auto kern = dj.greedy ? verify_mega_kernel<DT, true> : verify_mega_kernel<DT, false>;
  ws.fused = 1;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.stream = st; lc.attrs = at; lc.numAttrs = 1;
  lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(TS_THREADS); lc.dynamicSmemBytes = TS_SMEM;
  return cudaLaunchKernelEx(&lc, kern, dj, ws, cfg);
}

// Small batches of the plain modes: one launch, one thread-block cluster per sequence (cluster_small.cuh).
// Returns false when the shape is not eligible or the device cannot host one cluster (the caller takes the pipeline).
static int g_small_b = 0;       // largest batch that takes the cluster path (option "small_b"; 0 = never, the default:
                                // measured on B200 the one-launch path LOSES to the pipeline -- graph-replayed step
                                // 36 / 68 / 121 us vs 34 / 47 / 67 us at B = 1 / 32 / 64 -- its three phases serialise
                                // per sequence what the pipeline overlaps across sequences; kept opt-in and tested)
static int g_small_cl = 16;     // CTAs per cluster (option "small_cl": 8 or 16)
template <int DT>
static bool launch_cluster_small(const DecideJob& dj, HybridWs ws, const WsLayout& wl, void* workspace, int B, cudaStream_t st,
                                 cudaError_t& err) {
  const RowJob& rj = dj.rj;
  if (B > g_small_b || dj.gamma <= 0 || rj.top_k > 0 || rj.use_p || (dj.flags & SPECDEC_NGRAM) || g_no_fused_tail) return false;
  const int nseg = ((((rj.V + 7) >> 3) + 31) >> 5);
  int CL = g_small_cl >= 16 ? 16 : 8;
  if (nseg < CL) return false;  // tiny vocabularies: nothing to split
  auto kern = dj.greedy ? verify_cluster_kernel<DT, true> : verify_cluster_kernel<DT, false>;
  static int ok_dev[MAXDEV][2][2];  // per device / greedy / (CL == 16): 0 unknown, 1 usable, -1 not usable
  static size_t smem_dev[MAXDEV][2][2];
  static size_t attr_dev[MAXDEV][2];  // dynamic shared memory granted to the kernel so far (only ever raised)
  const int dv = cur_dev(), gi = dj.greedy ? 1 : 0;
  for (; CL >= 8; CL >>= 1) {
    const int ci = CL == 16 ? 1 : 0;
    const int spc = (nseg + CL - 1) / CL;
    const size_t smem = (size_t)spc * TF_SEG_BYTES;
    if (smem > 200 * 1024) return false;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.stream = st; lc.attrs = at; lc.numAttrs = 1;
    lc.gridDim = dim3((unsigned)B * CL); lc.blockDim = dim3(CS_T); lc.dynamicSmemBytes = smem;
    if (ok_dev[dv][gi][ci] == 0 || smem_dev[dv][gi][ci] != smem) {
      int ok = 1, ncl = 0;
      if (attr_dev[dv][gi] < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) ok = -1;
        else attr_dev[dv][gi] = smem;
      }
      if (ok == 1 && CL > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = -1;
      if (ok == 1 && (cudaOccupancyMaxActiveClusters(&ncl, kern, &lc) != cudaSuccess || ncl < 1)) ok = -1;
      cudaGetLastError();  // (a failed query must not poison the caller's next launch check)
      ok_dev[dv][gi][ci] = ok; smem_dev[dv][gi][ci] = smem;
    }
    if (ok_dev[dv][gi][ci] != 1) continue;
    ws.fused = 1;
    if ((err = cudaMemsetAsync((char*)workspace + wl.zero, 0, wl.zero_bytes_mega, st)) != cudaSuccess) return true;
    if (g_ev[0]) cudaEventRecord(g_ev[0], st);
    err = cudaLaunchKernelEx(&lc, kern, dj, ws, (float2*)((char*)workspace + wl.cpart), spc, CL);
    if (g_ev[1]) cudaEventRecord(g_ev[1], st);
    if (g_ev[2]) cudaEventRecord(g_ev[2], st);
    if (err == cudaSuccess) err = cudaGetLastError();
    return true;
  }
  return false;
}

template <int DT>
static cudaError_t launch_hybrid(DecideJob& dj, const WsLayout& wl, void* workspace, int B, cudaStream_t st) {
  const RowJob& rj = dj.rj;
  const bool masked = rj.top_k > 0 || rj.use_p;
  HybridWs ws = ws_pointers(wl, workspace, B, rj.R);
  {
    cudaError_t ce = cudaSuccess;
    if (launch_cluster_small<DT>(dj, ws, wl, workspace, B, st, ce)) return ce;
  }
  MegaCfg mcfg;
  int mgrid = 0;
  const bool mega = mega_config<DT>(dj, B, mcfg, mgrid);
  // (the self-validating exchange words of tail_slots_kernel / the megakernel live in the larger zeroed region)
  const bool slots_zero = !masked && dj.gamma > 0 && g_tail_slots && !g_no_fused_tail;
  cudaError_t e = cudaMemsetAsync((char*)workspace + wl.zero, 0, (mega || slots_zero) ? wl.zero_bytes_mega : wl.zero_bytes, st);
  if (e != cudaSuccess) return e;
  if (mega) {
    if (g_mega_dbg) {  // timeline: minima start at ~0ull, maxima at 0
      ws.dbg = (u64*)((char*)workspace + wl.dbg);
      cudaMemsetAsync(ws.dbg, 0, (size_t)(16 + B * 8 + 1024) * 8, st);
      cudaMemsetAsync(ws.dbg, 0xFF, 8, st);
      cudaMemsetAsync(ws.dbg + 3, 0xFF, 8, st);
      for (int b = 0; b < B; ++b) cudaMemsetAsync(ws.dbg + 16 + b * 8 + 2, 0xFF, 8, st);
    }
    if (g_ev[0]) cudaEventRecord(g_ev[0], st);
    if ((e = launch_mega<DT>(dj, ws, mcfg, mgrid, st)) != cudaSuccess) return e;
    if (g_ev[1]) cudaEventRecord(g_ev[1], st);
    if (g_ev[2]) cudaEventRecord(g_ev[2], st);
    return cudaGetLastError();
  }
  // Chunks of the batch pipelined on two streams: the HBM-bound row kernel of chunk i+1 overlaps the
  // latency/issue-bound exact tail of chunk i.  Results do not depend on the split.
  int C = g_chunks;
  if (masked || dj.gamma == 0 || B < 64 * C || C > 8 || DT == DT_F32) C = 1;  // (fp32 rows: LDG row kernel, no gain)
  AuxStream& ax = g_aux[cur_dev()];
  if (C > 1 && !ax.stream) {
    // (streams / events cannot be created while the caller's stream is being captured into a CUDA graph)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) C = 1;
  }
  if (C > 1 && !ax.stream) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&ax.stream, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
    for (int i = 0; i < 8; ++i)
      if ((e = cudaEventCreateWithFlags(&ax.ev_a[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&ax.ev_b, cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  if (g_ev[0]) cudaEventRecord(g_ev[0], st);
  if (C == 1) {
    if ((e = launch_phase_a<DT>(dj, ws, B, 3, st)) != cudaSuccess) return e;
    if (g_ev[1]) cudaEventRecord(g_ev[1], st);
    if ((e = launch_phase_b<DT>(dj, ws, B, st)) != cudaSuccess) return e;
  } else {
    for (int i = 0; i < C; ++i) {
      int b0 = (int)((long long)B * i / C), b1 = (int)((long long)B * (i + 1) / C);
      if (C == 2) {  // unequal halves: the exposed tail is the one of the last chunk
        const int cut = (int)((long long)B * g_chunk0_pct / 100);
        b0 = i ? cut : 0; b1 = i ? B : cut;
      }
      DecideJob d;
      HybridWs w;
      sub_job<DT>(dj, ws, b0, b1 - b0, i, d, w);
      if ((e = launch_phase_a<DT>(d, w, b1 - b0, i == 0 ? 3 : g_p1_ctas - 1, st, i > 0)) != cudaSuccess) return e;
      if (i == C - 1) {
        if (g_ev[1]) cudaEventRecord(g_ev[1], st);
        if ((e = launch_phase_b<DT>(d, w, b1 - b0, st)) != cudaSuccess) return e;
      } else {
        cudaEventRecord(ax.ev_a[i], st);
        cudaStreamWaitEvent(ax.stream, ax.ev_a[i], 0);
        if ((e = launch_phase_b<DT>(d, w, b1 - b0, ax.stream)) != cudaSuccess) return e;
      }
    }
    cudaEventRecord(ax.ev_b, ax.stream);
    cudaStreamWaitEvent(st, ax.ev_b, 0);
  }
  if (g_ev[2]) cudaEventRecord(g_ev[2], st);
  return cudaGetLastError();
}

#define DISPATCH_DT(dtype, ...)                                          \
  switch (dtype) {                                                       \
    case SPECDEC_F32: { constexpr int DT = DT_F32; __VA_ARGS__; } break;   \
    case SPECDEC_BF16: { constexpr int DT = DT_BF16; __VA_ARGS__; } break; \
    case SPECDEC_F16: { constexpr int DT = DT_F16; __VA_ARGS__; } break;   \
    default: return SPECDEC_ERR_DTYPE;                                   \
  }

}  // namespace specdec

using namespace specdec;

extern "C" {

int specdec_version(void) { return 100; }

const char* specdec_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SPECDEC_ERR_ARG: return "invalid argument";
    case SPECDEC_ERR_WORKSPACE: return "workspace too small";
    case SPECDEC_ERR_DTYPE: return "unsupported dtype";
    case SPECDEC_ERR_RANGE: return "value out of supported range";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

size_t specdec_workspace_bytes(int64_t rows) { return (size_t)(rows < 0 ? 0 : rows) * (sizeof(RowOut) + 16) + 256; }

size_t specdec_verify_workspace_bytes(int B, int gamma, int V) {
  if (B <= 0 || gamma < 0 || V <= 0) return 256;
  return ws_layout(B, gamma, V, (long long)B * (2 * gamma + 1)).total + 256;
}

int specdec_verify(const void* target_logits, const void* draft_logits, int dtype, const int64_t* draft_tokens,
                   const float* u_accept, const float* u_sample, uint64_t philox_seed, uint64_t philox_offset,
                   int64_t seq_id0, int B, int gamma, int V, int64_t stride_tb, int64_t stride_tg, int64_t stride_db,
                   int64_t stride_dg, float temperature, int top_k, float top_p, int sample_mode, int flags,
                   const int64_t* stop_tokens, int n_stop, int32_t* n_accepted, int64_t* next_token,
                   uint8_t* accept_mask, float* p_tok, float* q_tok, int32_t* first_stop, float* next_prob,
                   int32_t* packed, void* workspace, size_t workspace_bytes, specdec_stream_t stream) {
  std::lock_guard<std::mutex> api_lock(g_api_mu);
  if (B < 0 || gamma < 0 || gamma > 64 || !target_logits) return SPECDEC_ERR_ARG;
  if (B == 0) return 0;
  const bool ngram = flags & SPECDEC_NGRAM;
  if (!ngram && gamma > 0 && !draft_logits) return SPECDEC_ERR_ARG;
  if (gamma > 0 && (!draft_tokens || !accept_mask || !p_tok || !q_tok)) return SPECDEC_ERR_ARG;
  if (!n_accepted || !next_token || !first_stop) return SPECDEC_ERR_ARG;
  if (n_stop > 0 && !stop_tokens) return SPECDEC_ERR_ARG;
  if (sample_mode != SPECDEC_SAMPLE_GREEDY && sample_mode != SPECDEC_SAMPLE_INVCDF) return SPECDEC_ERR_ARG;
  const int nT = gamma + ((flags & SPECDEC_NO_BONUS) ? 0 : 1);
  const int nD = ngram ? 0 : gamma;
  if (nT + nD == 0) return SPECDEC_ERR_ARG;
  DecideJob dj;
  const long long R = (long long)B * (nT + nD);
  const WsLayout wl = ws_layout(B, gamma, V, R);
  if (workspace_bytes < wl.total) return SPECDEC_ERR_WORKSPACE;
  int rc = fill_rowjob(dj.rj, target_logits, draft_logits, stride_tb, stride_tg, stride_db, stride_dg, nT, nD, V,
                       temperature, top_k, top_p, R, workspace, workspace_bytes);
  if (rc) return rc;
  if (!ngram && !g_no_klist) dj.rj.klist = (int2*)((char*)workspace + wl.klist);
  if (!ngram) dj.rj.n_unres = ws_pointers(wl, workspace, B, R).ntasks + 14;  // (a word of the region zeroed per call)
  dj.draft_tokens = (const long long*)draft_tokens; dj.u_accept = u_accept; dj.u_sample = u_sample;
  dj.seed = philox_seed; dj.offset = philox_offset; dj.seq0 = seq_id0; dj.gamma = gamma;
  dj.offset_dev = nullptr;
  if (flags & SPECDEC_OFFSET_DEVICE) {
    if (!philox_offset) return SPECDEC_ERR_ARG;
    dj.offset_dev = (const u64*)(uintptr_t)philox_offset; dj.offset = 0;
  }
  dj.greedy = (sample_mode == SPECDEC_SAMPLE_GREEDY); dj.flags = flags;
  dj.stop = (const long long*)stop_tokens; dj.n_stop = n_stop;
  dj.n_acc = n_accepted; dj.next_tok = (long long*)next_token; dj.mask = accept_mask; dj.p_tok = p_tok;
  dj.q_tok = q_tok; dj.first_stop = first_stop; dj.next_prob = next_prob; dj.packed = packed;
  dj.lane_sample = 0x10000;
  cudaStream_t st = (cudaStream_t)stream;
  const bool masked_mode = dj.rj.top_k > 0 || dj.rj.use_p;
  if (ngram && dj.greedy && !masked_mode && !g_no_fast_ngram) {
    // greedy n-gram verify: arg-max per row from the fast row kernel, no exact pass (hybrid.cuh)
    HybridWs ws = ws_pointers(wl, workspace, B, dj.rj.R);
    if (g_ev[0]) cudaEventRecord(g_ev[0], st);
    DISPATCH_DT(dtype, {
      if (!ngram_argmax_tma<DT>(dj.rj, st))  // 16-bit aligned rows: the TMA row pipeline with an arg-max epilogue
        rowfast_argmax_kernel<DT><<<(unsigned)dj.rj.R, FT, 0, st>>>(dj.rj);
      if (g_ev[1]) cudaEventRecord(g_ev[1], st);
      ngram_greedy_decide_kernel<DT><<<B, PT, 0, st>>>(dj, ws);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
    });
    if (g_ev[2]) cudaEventRecord(g_ev[2], st);
    return 0;
  }
  if (ngram) {  // per-position sample(p_i) comparisons: exact statistics for every row, one CTA per sequence
    if (g_ev[0]) cudaEventRecord(g_ev[0], st);
    DISPATCH_DT(dtype, {
      cudaError_t e = launch_rowstats<DT>(dj.rj, st);
      if (e != cudaSuccess) return (int)e;
      if (g_ev[1]) cudaEventRecord(g_ev[1], st);
      decide_kernel<DT><<<B, NT, 0, st>>>(dj);
      e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
    });
    if (g_ev[2]) cudaEventRecord(g_ev[2], st);
    return 0;
  }
  DISPATCH_DT(dtype, {
    cudaError_t e = launch_hybrid<DT>(dj, wl, workspace, B, st);
    if (e != cudaSuccess) return (int)e;
  });
  return 0;
}

int specdec_set_option(const char* name, int value) {
  std::lock_guard<std::mutex> api_lock(g_api_mu);
  if (!name) return SPECDEC_ERR_ARG;
  if (!strcmp(name, "reset")) {  // every option back to its default (tests call this after each case)
    g_force_ldg = 0; g_chunks = 2; g_chunk0_pct = 50; g_p1_ctas = 3; g_tf_ch = TF_CH_DEFAULT; g_no_fast_nucleus = 0;
    g_no_hist_nucleus = 0; g_no_tma_nucleus = 1; g_no_fast_ngram = 0; g_no_fused_tail = 0; g_no_pdl = 0; g_tma_ngram = 1;
    g_no_klist = 0; g_tail_slots = 1; g_tf_balance = 1; g_static_rows = 0; g_split_lists = 0; g_small_b = 0; g_small_cl = 16; g_no_rowsel = 0; g_rowsel_probe = 0; g_mega = 0; g_mega_r = 2; g_mega_unit = 4; g_mega_spc = 24; g_mega_keep_l2 = 1; g_mega_dbg = 0;
    return 0;
  }
  if (!strcmp(name, "force_ldg")) { g_force_ldg = value; return 0; }
  if (!strcmp(name, "no_overlap")) { g_chunks = value ? 1 : 2; return 0; }
  if (!strcmp(name, "chunks")) { if (value < 0 || value > 8) return SPECDEC_ERR_ARG; g_chunks = value ? value : 2; return 0; }  // 0 = default
  if (!strcmp(name, "chunk0_pct")) { if (value < 10 || value > 90) return SPECDEC_ERR_ARG; g_chunk0_pct = value; return 0; }
  if (!strcmp(name, "p1_ctas")) { if (value < 1 || value > 4) return SPECDEC_ERR_ARG; g_p1_ctas = value; return 0; }
  if (!strcmp(name, "tf_ch")) { if (value < 2 || value > 64) return SPECDEC_ERR_ARG; g_tf_ch = value; return 0; }
  if (!strcmp(name, "mega")) { g_mega = value; return 0; }
  if (!strcmp(name, "mega_r")) { if (value < 1 || value > 3) return SPECDEC_ERR_ARG; g_mega_r = value; return 0; }
  if (!strcmp(name, "mega_unit")) { if (value < 1 || value > 64) return SPECDEC_ERR_ARG; g_mega_unit = value; return 0; }
  if (!strcmp(name, "mega_spc")) { if (value < MG_MIN_SPC || value > TS_SMEM / TF_SEG_BYTES) return SPECDEC_ERR_ARG; g_mega_spc = value; return 0; }
  if (!strcmp(name, "mega_keep_l2")) { g_mega_keep_l2 = value; return 0; }
  if (!strcmp(name, "mega_dbg")) { g_mega_dbg = value; return 0; }
  if (!strcmp(name, "no_fast_nucleus")) { g_no_fast_nucleus = value; return 0; }
  if (!strcmp(name, "no_klist")) { g_no_klist = value; return 0; }
  if (!strcmp(name, "small_b")) { g_small_b = value; return 0; }
  if (!strcmp(name, "tail_slots")) { g_tail_slots = value; return 0; }
  if (!strcmp(name, "tf_balance")) { g_tf_balance = value; return 0; }
  if (!strcmp(name, "static_rows")) { g_static_rows = value; return 0; }
  if (!strcmp(name, "split_lists")) { g_split_lists = value; return 0; }
  if (!strcmp(name, "small_cl")) { g_small_cl = value; return 0; }
  if (!strcmp(name, "no_rowsel")) { g_no_rowsel = value; return 0; }
  if (!strcmp(name, "rowsel_probe")) { g_rowsel_probe = value; return 0; }
  if (!strcmp(name, "no_hist_nucleus")) { g_no_hist_nucleus = value; return 0; }
  if (!strcmp(name, "no_tma_nucleus")) { g_no_tma_nucleus = value; return 0; }
  if (!strcmp(name, "no_fast_ngram")) { g_no_fast_ngram = value; return 0; }
  if (!strcmp(name, "no_fused_tail")) { g_no_fused_tail = value; return 0; }
  if (!strcmp(name, "no_pdl")) { g_no_pdl = value; return 0; }
  if (!strcmp(name, "tma_ngram")) { g_tma_ngram = value; return 0; }
  return SPECDEC_ERR_ARG;
}

int specdec_debug_stats(unsigned long long* out16, int reset) {
  if (!out16) return SPECDEC_ERR_ARG;
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_nh_stats, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return (int)e;
  if (reset) {
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(g_nh_stats, z, sizeof(z));
  }
  return (int)e;
}

int specdec_debug_timeline(const void* workspace, int B, int gamma, int V, unsigned long long* host_out, int n) {
  if (!workspace || !host_out || B <= 0) return SPECDEC_ERR_ARG;
  const WsLayout wl = ws_layout(B, gamma, V, (long long)B * (2 * gamma + 1));
  const int have = 16 + B * 8 + 1024;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemcpy(host_out, (const char*)workspace + wl.dbg, sizeof(unsigned long long) * (size_t)(n < have ? n : have), cudaMemcpyDeviceToHost);
}

int specdec_set_profile_events(void* ev_start, void* ev_mid, void* ev_end) {
  std::lock_guard<std::mutex> api_lock(g_api_mu);
  g_ev[0] = (cudaEvent_t)ev_start; g_ev[1] = (cudaEvent_t)ev_mid; g_ev[2] = (cudaEvent_t)ev_end;
  return 0;
}

int specdec_process_probs(const void* logits, int dtype, int64_t rows, int V, int64_t stride, float temperature,
                          int top_k, float top_p, float* probs, float* row_stats, void* workspace,
                          size_t workspace_bytes, specdec_stream_t stream) {
  std::lock_guard<std::mutex> api_lock(g_api_mu);
  if (rows < 0 || !logits) return SPECDEC_ERR_ARG;
  if (rows == 0) return 0;
  RowJob rj;
  int rc = fill_rowjob(rj, logits, nullptr, stride, 0, 0, 0, 1, 0, V, temperature, top_k, top_p, rows, workspace,
                       workspace_bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, {
    cudaError_t e = launch_rowstats<DT>(rj, st);
    if (e != cudaSuccess) return (int)e;
    const int grid = (int)(rows < 4LL * num_sms() ? rows : 4LL * num_sms());
    probs_kernel<DT><<<grid, NT, 0, st>>>(rj, probs, row_stats);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  });
  return 0;
}

int specdec_sample_rows(const void* logits, int dtype, int64_t rows, int V, int64_t stride, float temperature,
                        int top_k, float top_p, int sample_mode, const float* u, uint64_t philox_seed,
                        uint64_t philox_offset, int64_t seq_id0, int lane_id, int64_t* tok, float* ptok,
                        void* workspace, size_t workspace_bytes, specdec_stream_t stream) {
  std::lock_guard<std::mutex> api_lock(g_api_mu);
  if (rows < 0 || !logits || !tok) return SPECDEC_ERR_ARG;
  if (rows == 0) return 0;
  if (rows > 0x7fffffff) return SPECDEC_ERR_RANGE;
  if (sample_mode != SPECDEC_SAMPLE_GREEDY && sample_mode != SPECDEC_SAMPLE_INVCDF) return SPECDEC_ERR_ARG;
  // a verify step with gamma = 0: every sequence "accepts all" and draws its bonus token
  const WsLayout wl = ws_layout(rows, 0, V, rows);
  const size_t need = wl.total + (size_t)rows * (sizeof(int) * 2);
  if (workspace_bytes < need) return SPECDEC_ERR_WORKSPACE;
  DecideJob dj;
  int rc = fill_rowjob(dj.rj, logits, nullptr, stride, 0, 0, 0, 1, 0, V, temperature, top_k, top_p, rows, workspace,
                       workspace_bytes);
  if (rc) return rc;
  int* scratch = (int*)((char*)workspace + wl.total);
  if (!g_no_klist) dj.rj.klist = (int2*)((char*)workspace + wl.klist);
  dj.rj.n_unres = ws_pointers(wl, workspace, rows, rows).ntasks + 14;
  dj.draft_tokens = nullptr; dj.u_accept = nullptr; dj.u_sample = u;
  dj.seed = philox_seed; dj.offset = philox_offset; dj.offset_dev = nullptr; dj.seq0 = seq_id0; dj.gamma = 0;
  dj.greedy = (sample_mode == SPECDEC_SAMPLE_GREEDY); dj.flags = 0; dj.stop = nullptr; dj.n_stop = 0;
  dj.n_acc = scratch; dj.first_stop = scratch + rows; dj.next_tok = (long long*)tok; dj.mask = nullptr;
  dj.p_tok = nullptr; dj.q_tok = nullptr; dj.next_prob = ptok; dj.packed = nullptr;
  if (lane_id & SPECDEC_LANE_OFFSET_DEVICE) {
    if (!philox_offset) return SPECDEC_ERR_ARG;
    lane_id &= ~SPECDEC_LANE_OFFSET_DEVICE;
    dj.offset_dev = (const u64*)(uintptr_t)philox_offset; dj.offset = 0;
  }
  dj.lane_sample = 0x20000 + lane_id;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_DT(dtype, {
    cudaError_t e = launch_hybrid<DT>(dj, wl, workspace, (int)rows, st);
    if (e != cudaSuccess) return (int)e;
  });
  return 0;
}

size_t specdec_sample_rows_workspace_bytes(int64_t rows, int V) {
  if (rows <= 0 || V <= 0) return 256;
  return ws_layout(rows, 0, V, rows).total + (size_t)rows * 8 + 256;
}

int specdec_sample_probs(const float* probs, int64_t rows, int V, int sample_mode, const float* u, int64_t* tok,
                         specdec_stream_t stream) {
  if (rows < 0 || !probs || !tok) return SPECDEC_ERR_ARG;
  if (rows == 0) return 0;
  if (V <= 0 || V > MAXPART * 256) return SPECDEC_ERR_RANGE;
  const int greedy = (sample_mode == SPECDEC_SAMPLE_GREEDY);
  if (!greedy && !u) return SPECDEC_ERR_ARG;
  const int grid = (int)(rows < 4LL * num_sms() ? rows : 4LL * num_sms());
  sample_probs_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(probs, rows, V, greedy, u, (long long*)tok);
  return (int)cudaGetLastError();
}

int specdec_philox_uniform(uint64_t seed, uint64_t offset, int64_t seq_id0, int B, int gamma, float* u_accept,
                           float* u_sample, specdec_stream_t stream) {
  if (B < 0 || gamma < 0) return SPECDEC_ERR_ARG;
  const int n = B * (gamma + 1);
  if (n == 0) return 0;
  philox_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(seed, offset, seq_id0, B, gamma, u_accept, u_sample);
  return (int)cudaGetLastError();
}

}  // extern "C"
