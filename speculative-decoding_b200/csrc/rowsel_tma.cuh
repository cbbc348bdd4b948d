// rowsel_tma.cuh -- masked modes (top-k, top-k + top-p, top-p on rows with a small nucleus) at HBM speed: the exact
// kept set and normaliser of every row from ONE streamed pass (included by verify.cu inside namespace specdec).
//
// rowstats_kernel / nucleus_fast_kernel read a row with plain loads (45 % of the HBM peak), then gather the candidates.
// Here the row is streamed by the TMA bulk-copy pipeline of rowfast_tma.cuh, warp-specialised three ways:
//   producer warp   cp.async.bulk HBM -> 4 x 16 KB ring, mbarrier transaction counts (as rowfast_tma_kernel)
//   8 consumer warps  per stage: maximum on the packed 16-bit words (HMNMX2), kept as the three largest stage maxima of
//                   the thread's slice with the stages of the first two (8 branch-free instructions per stage),
//                   [top-p: online MUFU T=1 mass];  per row: the threshold among the 256 slice maxima (slice of thread
//                   t = vectors v = t mod 256), found by one warp -- top-k: the k-th largest (>= k elements above it);
//                   top-p: the largest value whose slice maxima alone out-weigh top_p * S1 -- then every thread whose
//                   maximum reaches the threshold looks at its one or two hot STAGES again (4 vectors each, from L2: 2
//                   CTAs per SM keep 76 MB of rows in flight) and appends its elements above it to a candidate buffer
//   selector warp   exact selection among the candidates (select_cut_group, integer arithmetic), RowOut record and
//                   kept-token list -- concurrently with the consumers streaming the next row (two candidate buffers)
// Rows it cannot resolve (candidate overflow, ambiguous top-p bracket, flat top-p rows) stay unflagged and are redone
// by nucleus_hist_kernel / rowstats_kernel.  Results are bit-identical to those kernels: the same selection code runs
// on a superset of the kept set.
#pragma once

constexpr int RS2_CONSUMERS = 256;
constexpr int RS2_SELECTORS = 1;  // selector warps (2: one per candidate buffer -- measured slower: 352 vs 310 us on top-k + top-p)
constexpr int RS2_THREADS = RS2_CONSUMERS + 32 + 32 * RS2_SELECTORS;  // + producer warp + selector warp(s)
#ifndef RS2_STAGES_V
#define RS2_STAGES_V 4
#endif
#ifndef RS2_CAP_V
#define RS2_CAP_V 1024
#endif
constexpr int RS2_STAGES = RS2_STAGES_V;
constexpr int RS2_STAGE_BYTES = 16384;
constexpr int RS2_CAP = RS2_CAP_V;  // candidates per buffer (<= WARP_SELECT_MAX: one warp selects)
constexpr size_t RS2_CAND_BYTES = (size_t)RS2_CAP * (sizeof(float) + sizeof(int) + sizeof(u64));
constexpr size_t RS2_SMEM = (size_t)RS2_STAGES * RS2_STAGE_BYTES + 2 * RS2_CAND_BYTES;

struct Rs2Meta {  // consumers -> selector, one per candidate buffer
  long long r;
  float m, S1f;
  int n, ok;
};

template <int DT, bool HK, bool HP>
__global__ void __launch_bounds__(RS2_THREADS, 2) rowsel_tma_kernel(RowJob job) {
  extern __shared__ __align__(128) unsigned char rs2_dyn[];
  __shared__ __align__(8) unsigned long long full_bar[RS2_STAGES], empty_bar[RS2_STAGES];
  __shared__ float sh_f[2][8], sh_g[2][8];
  __shared__ int s_count[2];
  __shared__ Rs2Meta meta[2];
  __shared__ volatile int buf_state[2];  // 0: free (consumers may fill), 1: ready for the selector
  unsigned char* ring = rs2_dyn;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned row_bytes = (unsigned)job.V * 2u;
  const int nst = (int)((row_bytes + RS2_STAGE_BYTES - 1) / RS2_STAGE_BYTES);
  const int V = job.V, NV = (V + 7) >> 3;
  const float c = job.c, c1 = job.c1;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < RS2_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], RS2_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    buf_state[0] = 0; buf_state[1] = 0;
  }
  __syncthreads();
  if (warp == RS2_CONSUMERS / 32) {
    // ---------------- producer warp ----------------
    if (lane == 0) {
      int stage = 0;
      unsigned phase = 0;
      const unsigned long long policy = l2_evict_normal_policy();  // the hot slices are read again within microseconds
      for (long long r = blockIdx.x; r < job.R; r += gridDim.x) {
        const char* base = (const char*)row_ptr<DT>(job, r);
        for (int k = 0; k < nst; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const unsigned off = (unsigned)k * RS2_STAGE_BYTES;
          const unsigned nb = min((unsigned)RS2_STAGE_BYTES, row_bytes - off);
          mbar_expect_tx(&full_bar[stage], nb);
          tma_bulk_g2s(ring + stage * RS2_STAGE_BYTES, base + off, nb, &full_bar[stage], policy);
          if (++stage == RS2_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }
  if (warp >= RS2_CONSUMERS / 32 + 1) {
    // ---------------- selector warps: warp s owns candidate buffer s, i.e. every second row of this CTA (the exact
    // selection of a top-k + top-p row takes one warp longer than the consumers need to stream a row)
    int par = warp - (RS2_CONSUMERS / 32 + 1);
    for (long long r = blockIdx.x + (long long)par * gridDim.x; r < job.R; r += (long long)RS2_SELECTORS * gridDim.x,
                   par = (RS2_SELECTORS == 1) ? (par ^ 1) : par) {
      if (lane == 0)
        while (buf_state[par] != 1) __nanosleep(64);
      __syncwarp();
      __threadfence_block();
      const Rs2Meta mt = meta[par];
      float* cz = (float*)(rs2_dyn + (size_t)RS2_STAGES * RS2_STAGE_BYTES + par * RS2_CAND_BYTES);
      int* cj = (int*)(cz + RS2_CAP);
      u64* cw = (u64*)(cj + RS2_CAP);
      const float m = mt.m, mc = __fmul_rn(m, c), mc1 = __fmul_rn(m, c1);
      const int n = mt.n;
      bool ok = mt.ok && n > 0 && n <= RS2_CAP && (!HK || n >= job.top_k) && !job.pre_stats;  // (pre_stats: timing probe)
      float cut = -INFINITY;
      int jcut = V, amb = 0, kc = 0;
      u64 Sfix = 0;
      if (ok) {
        const Grp<false> gp{nullptr, nullptr};
        int mval = 0;
        if (HK) {
          select_cut_group<false>(gp, cz, cj, cw, n, V, job.top_k, HP ? 1 : 0, job.tpq, 0ull, 0ull, 0ull, c, mc, c1, mc1, cut, jcut,
                                  Sfix, 0, 0, nullptr, &mval);
        } else {
          // S1 in 2^-40 fixed point, bracketed by the MUFU error (1e-4 is ~50x the observed 2e-6): if both ends of the
          // bracket select the same kept set it is the exact one (and the candidates hold the whole nucleus)
          const double S1d = (double)mt.S1f * 1099511627776.0;
          const u64 thr_lo = scale_q32((u64)(S1d * (1.0 - 1e-4)), job.tpq), thr_hi = scale_q32((u64)(S1d * (1.0 + 1e-4)), job.tpq);
          amb = 1;
          select_cut_group<false>(gp, cz, cj, cw, n, V, 0, 1, job.tpq, 0ull, 0ull, 0ull, c, mc, c1, mc1, cut, jcut, Sfix, thr_lo,
                                  thr_hi, &amb, &mval);
        }
        ok = (amb == 0);
        if (ok && job.klist) kc = write_kept_list(cz, cj, mval, cut, jcut, job.klist + mt.r * KL_MAX);
      }
      if (lane == 0) {
        RowOut o;
        o.m = m; o.mc = mc; o.cut = -INFINITY; o.jcut = V; o.flags = 0; o.Sfix = 0;
        o.inv = mt.S1f;  // unresolved rows: the MUFU T=1 mass (relative to the max) for nucleus_hist_kernel
        if (ok) {
          o.cut = cut; o.jcut = jcut; o.Sfix = Sfix; o.flags = 1 | (kc << 8);
          o.inv = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(Sfix), 0x1p-40f));
        }
        job.out[mt.r] = o;
        if (!ok && job.n_unres) atomicAdd(job.n_unres, 1);
      }
      __syncwarp();
      __threadfence_block();
      if (lane == 0) buf_state[par] = 0;
    }
    return;
  }
  // ---------------- consumer warps ----------------
  __shared__ float tm_sh[RS2_CONSUMERS];
  __shared__ unsigned s_K;
  __shared__ int s_flat, s_ovf[2];
  int stage = 0, par = 0;
  unsigned phase = 0;
  for (long long r = blockIdx.x; r < job.R; r += gridDim.x, par ^= 1) {
    // m1 >= m2 >= m3 >= m4: the four largest STAGE maxima of my slice (k1, k2, k3: the stages of the first three).
    // Every element outside stages k1..k3 is <= m4, so after the row only those stages (4 vectors each) have to be
    // looked at again to list all my elements above a threshold th > m4.  (With two tracked stages a third of the
    // top-50 rows overflowed -- 50 candidates in 256 slices put three into one slice that often; with three it is 2 %.)
    float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY, m4 = -INFINITY, s = 0.0f;
    int k1 = 0, k2 = 0, k3 = 0;
    for (int k = 0; k < nst; ++k) {
      mbar_wait(&full_bar[stage], phase);
      const unsigned off = (unsigned)k * RS2_STAGE_BYTES;
      const int nvec = (int)(min((unsigned)RS2_STAGE_BYTES, row_bytes - off) >> 4);
      const uint4* sp = reinterpret_cast<const uint4*>(ring + stage * RS2_STAGE_BYTES);
      constexpr int VPT = RS2_STAGE_BYTES / 16 / RS2_CONSUMERS;
      uint4 a[VPT];
#pragma unroll
      for (int q = 0; q < VPT; ++q)
        if (tid + q * RS2_CONSUMERS < nvec) a[q] = sp[tid + q * RS2_CONSUMERS];
      unsigned pm = (DT == DT_BF16) ? 0xFF80FF80u : 0xFC00FC00u;  // (-inf, -inf)
#pragma unroll
      for (int q = 0; q < VPT; ++q)
        if (tid + q * RS2_CONSUMERS < nvec) pm = packed_max4<DT>(a[q], pm);
      const float vm = packed_max_to_float<DT>(pm);
      if (HP && !HK) {  // pure top-p: the T = 1 mass of the row (MUFU, 1e-6) for the nucleus threshold
        if (vm > m1) s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m1, vm), c1)));
        const float mm = fmaxf(m1, vm);
        const float nmc = (mm > -INFINITY) ? -__fmul_rn(mm, c1) : 0.0f;
        const float2 c2 = make_float2(c1, c1), nmc2 = make_float2(nmc, nmc);
        float2 acc = make_float2(s, 0.0f);
#pragma unroll
        for (int q = 0; q < VPT; ++q)
          if (tid + q * RS2_CONSUMERS < nvec) accumulate16<DT>(a[q], c2, nmc2, acc);
        s = __fadd_rn(acc.x, acc.y);
      }
      {  // branch-free insertion of vm into (m1, m2, m3, m4)
        const bool g1 = vm > m1, g2 = vm > m2, g3 = vm > m3;
        m4 = g3 ? m3 : fmaxf(m4, vm);
        k3 = g2 ? k2 : (g3 ? k : k3);
        m3 = g2 ? m2 : (g3 ? vm : m3);
        k2 = g1 ? k1 : (g2 ? k : k2);
        m2 = g1 ? m1 : (g2 ? vm : m2);
        k1 = g1 ? k : k1;
        m1 = g1 ? vm : m1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == RS2_STAGES) { stage = 0; phase ^= 1u; }
    }
    // ---- row epilogue among the 256 consumers (named barrier 1); the producer keeps prefetching the next row
    const float m = m1;
    const float wm = warp_max_f(m);
    float ws_ = 0.0f;
    if (HP && !HK) {
      const float resc = (m > -INFINITY) ? ex2_approx(__fmul_rn(__fsub_rn(m, wm), c1)) : 0.0f;
      ws_ = (m > -INFINITY) ? __fmul_rn(s, resc) : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ws_ += __shfl_xor_sync(0xffffffffu, ws_, o);
    }
    const int ep = par;  // scratch double-buffered by row parity
    if (lane == 0) { sh_f[ep][warp] = wm; sh_g[ep][warp] = ws_; }
    tm_sh[tid] = m;
    // the candidate buffer of this parity must have been consumed by the selector (two rows ago)
    if (tid == 0) {
      while (buf_state[par] != 0) __nanosleep(64);
      s_count[par] = 0; s_ovf[par] = 0;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RS2_CONSUMERS) : "memory");
    float M = sh_f[ep][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) M = fmaxf(M, sh_f[ep][w]);
    float S1f = 0.0f;
    if (HP && !HK) {
#pragma unroll
      for (int w = 0; w < 8; ++w)
        S1f += (sh_f[ep][w] > -INFINITY) ? __fmul_rn(sh_g[ep][w], ex2_approx(__fmul_rn(__fsub_rn(sh_f[ep][w], M), c1))) : 0.0f;
    }
    // ---- threshold among the 256 slice maxima, by ONE warp: the largest key K with  sum_{key >= K} weight > t0
    //      (top-k: weight 1, t0 = k - 1/2;  top-p: weight = MUFU mass of the slice maximum, t0 = top_p * S1 * 1.001: the
    //      slice maxima above th alone out-weigh the nucleus, so the nucleus lies above th -- verified with exact masses
    //      by the selector's bracket test)
    if (warp == 0) {
      float tv[8], wg[8];
      unsigned ky[8];
      const unsigned refkey = fkey(M);
      unsigned vary = 0;
      float tot = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        tv[i] = tm_sh[i * 32 + lane];
        ky[i] = fkey(tv[i]);
        vary |= ky[i] ^ refkey;
        wg[i] = HK ? 1.0f : ((tv[i] > -INFINITY) ? ex2_approx(__fmul_rn(__fsub_rn(tv[i], M), c1)) : 0.0f);
        tot += wg[i];
      }
      vary = __reduce_or_sync(0xffffffffu, vary);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      const float t0 = HK ? (float)job.top_k - 0.5f : (float)((double)job.tpq * (1.0 / 4294967296.0)) * S1f * 1.001f;
      unsigned K = refkey & ~vary;
      for (int bit = 31; bit >= 0; --bit) {
        if (!((vary >> bit) & 1u)) continue;
        const unsigned tr = K | (1u << bit);
        float sw = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sw += (ky[i] >= tr) ? wg[i] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, o);
        if (sw > t0) K = tr;
      }
      if (lane == 0) { s_K = K; s_flat = (tot > t0) ? 0 : 1; }  // flat: the slice maxima do not reach the nucleus mass
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RS2_CONSUMERS) : "memory");
    const unsigned K = s_K;
    bool ok = (M > -INFINITY) && (M < INFINITY) && !s_flat && (!HK || job.top_k <= RS2_CONSUMERS);
    const float th = fkey_inv(K);
    float* cz = (float*)(rs2_dyn + (size_t)RS2_STAGES * RS2_STAGE_BYTES + par * RS2_CAND_BYTES);
    int* cj = (int*)(cz + RS2_CAP);
    if (ok && m1 >= th) {
      // my slice holds candidates: look at stage k1 (and k2, k3 if their maxima reach th) again -- from L2, the row
      // was streamed microseconds ago.  More than three hot stages (2 % of the top-50 rows): this thread looks at ALL
      // its stages again -- a dozen microseconds for that row's epilogue, instead of leaving the whole row to the
      // exact fallback kernel (one 512-thread CTA sweeping it twice: 57 us on the critical path of every step)
      const bool all_stages = (m4 >= th);
      const char* row = (const char*)row_ptr<DT>(job, r);
      const int nlook = all_stages ? nst : ((m3 >= th) ? 3 : ((m2 >= th) ? 2 : 1));
      constexpr int VPT = RS2_STAGE_BYTES / 16 / RS2_CONSUMERS;
      for (int which = 0; which < nlook; ++which) {
        const int kk = all_stages ? which : (which == 0 ? k1 : (which == 1 ? k2 : k3));
        uint4 a[VPT];
        int vv[VPT];
#pragma unroll
        for (int q = 0; q < VPT; ++q) {
          vv[q] = kk * (RS2_STAGE_BYTES / 16) + tid + q * RS2_CONSUMERS;
          a[q] = (vv[q] < NV) ? __ldg(reinterpret_cast<const uint4*>(row) + vv[q]) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < VPT; ++q) {
          if (vv[q] >= NV) continue;
          Raw8<DT> rw; rw.a = a[q];
          float x[8];
          unpack8<DT>(rw, x);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (x[e] >= th && vv[q] * 8 + e < V) {
              const int pos = atomicAdd(&s_count[par], 1);
              if (pos < RS2_CAP) { cz[pos] = x[e]; cj[pos] = vv[q] * 8 + e; }
            }
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RS2_CONSUMERS) : "memory");
    if (tid == 0) {
      meta[par].r = r; meta[par].m = M; meta[par].S1f = S1f; meta[par].n = s_count[par];
      meta[par].ok = (ok && !s_ovf[par]) ? 1 : 0;
      __threadfence_block();
      buf_state[par] = 1;
    }
  }
}
