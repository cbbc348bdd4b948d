// tail_fused.cuh -- everything after the row kernel + plan in ONE launch (included by hybrid.cuh).
//
// The exact part of a verify step is instruction-issue bound: exact_rows_kernel evaluated the canonical
// exp of the deciding row pair (P_n, Q_n) for the normalisers, and sample_partial_kernel evaluated the
// same exponentials AGAIN for the residual max(0, P - Q), because the residual needs 1/S of the whole
// row first.  Here TF_CH CTAs share one sequence: each evaluates its slice of the row pair once, keeps
// the weights e_p, e_q (fp32) in shared memory (64 KB per CTA at V = 128256), publishes its partial
// canonical sums with u64 atomics and waits until the TF_CH partial sums of the sequence have arrived;
// the residual partial sums then come from shared memory (7 instead of 36 instructions per pair).
//
// Waiting on sibling CTAs needs a forward-progress guarantee: the logical CTA id is a ticket taken
// from an atomic counter at CTA start, so a running CTA only ever waits for CTAs with tickets in its
// own group of TF_CH -- all of which are running or are the next to be dispatched (TF_CH <= resident
// CTAs).  Rare paths: sequences with ambiguous accept tests first get the exact sums of those
// positions (same CTAs, no caching), CTA 0 of the group decides, the others wait for its flag.
// Results are bit-identical to the exact_rows/sample_partial pipeline (same integers are summed).
#pragma once

constexpr int TF_CH_DEFAULT = 20;  // CTAs per sequence (52 KB of weights per CTA at V = 128256: 4 CTAs / SM)
constexpr int TF_T = 256;   // threads per CTA (== PT: partial_loop assumes it)
constexpr int TF_SEG_BYTES = 2048;  // one 256-pair segment of (e_p, e_q)

// row k (target rows first, then drafter rows) of sequence b, without row_ptr's 64-bit division
template <int DT>
__device__ __forceinline__ const void* seq_row_ptr(const RowJob& job, int b, int k) {
  const size_t es = (DT == DT_F32) ? 4 : 2;
  if (k < job.nT) return (const char*)job.tgt + (size_t)((long long)b * job.tsb + (long long)k * job.tsg) * es;
  return (const char*)job.drf + (size_t)((long long)b * job.dsb + (long long)(k - job.nT) * job.dsg) * es;
}
// two u64 block sums with one pair of barriers; `sh` holds 66 entries
__device__ __forceinline__ void block_sum2_u64(u64& a, u64& b, u64* sh) {
  a = warp_sum_u64(a);
  b = warp_sum_u64(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (cta_nthreads() + 31) >> 5;
  if (lane == 0) { sh[w] = a; sh[33 + w] = b; }
  cta_sync();
  u64 t = (lane < nw) ? sh[lane] : 0ull, u = (lane < nw) ? sh[33 + lane] : 0ull;
  a = warp_sum_u64(t);
  b = warp_sum_u64(u);
  cta_sync();
}

// Canonical sums of segments [s0, s1) of a row pair; CACHE: also keep the weights, SCALED by 2^40 (cweight2_s40), in
// shared memory as float4 {e_p[2k], e_p[2k+1], e_q[2k], e_q[2k+1]} at [(seg - s0) * 128 + k * 32 + lane] (conflict-free).
template <int DT, bool CACHE, int NS_ = 0>
__device__ __forceinline__ void pair_sums(const void* prow, const void* qrow, bool pal, bool qal, int V, float c,
                                          float mcp, float mcq, int s0, int s1, float4* ecache, u64& sp, u64& sq) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int NV = (V + 7) >> 3;
  const float2 c2 = make_float2(c, c), nmp2 = make_float2(-mcp, -mcp), nmq2 = make_float2(-mcq, -mcq);
  constexpr int WPB = TF_T / 32;
  constexpr int NS = NS_ ? NS_ : ((DT == DT_F32) ? 2 : 4);  // segments in flight per warp (raw, still packed loads)
  for (int seg = s0 + w; seg < s1; seg += NS * WPB) {
    Raw8<DT> rp[NS], rq[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int vv = min((seg + q * WPB) * 32 + lane, NV - 1);
      rp[q] = load_raw8<DT>(prow, vv, V, pal);
      rq[q] = load_raw8<DT>(qrow, vv, V, qal);
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int sg = seg + q * WPB, vv = sg * 32 + lane;
      if (sg < s1) {  // warp-uniform
        float4* dst = CACHE ? ecache + (size_t)(sg - s0) * 128 + lane : nullptr;
        if (vv < NV) {
          float xp[8], xq[8];
          unpack8<DT>(rp[q], xp);
          unpack8<DT>(rq[q], xq);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // two neighbouring tokens of one row as one fp32x2 pair, weights * 2^40
            const float2 ep = cweight2_s40(make_float2(xp[2 * k], xp[2 * k + 1]), c2, nmp2);
            const float2 eq = cweight2_s40(make_float2(xq[2 * k], xq[2 * k + 1]), c2, nmq2);
            sp += __float2ull_rz(ep.x) + __float2ull_rz(ep.y);  // == fix40(e): e * 2^40 is exact
            sq += __float2ull_rz(eq.x) + __float2ull_rz(eq.y);
            if (CACHE) dst[k * 32] = make_float4(ep.x, ep.y, eq.x, eq.y);
          }
        } else if (CACHE) {
#pragma unroll
          for (int k = 0; k < 4; ++k) dst[k * 32] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
      }
    }
  }
}

struct TailSh {  // static shared memory of one tail group
  u64 sh64[66];
  float shf[33];
  int shi[33];
  long long s_res;
  int sh_last, sh_ok;
  u64 sh_S[2];
};

// Everything after the plan for slice `ch` (of TF_CH) of sequence b.  Called by all threads of the CTA's compute group
// (cta_nthreads() threads, cta_sync() barriers; work loops use the first TF_T of them).  Returns false iff a bounded
// inter-CTA wait gave up (ws.abort is then set and the caller leaves the kernel).
template <int DT, bool GREEDY>
__device__ __forceinline__ bool tail_item(const DecideJob& job, const HybridWs& ws, const int b, const int ch, const int TF_CH,
                                          const int segs_per_cta, float4* ecache, TailSh& sh, const int seq_tasks,
                                          const bool have_rec, int4 rec, int mcq_bits) {
  const RowJob& rj = job.rj;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = job.gamma, V = rj.V, rps = rj.nT + rj.nD;
  const int NV = (V + 7) >> 3, nseg = (NV + 31) >> 5;
  const int s0 = min(nseg, ch * segs_per_cta), s1 = min(nseg, s0 + segs_per_cta);
  const float c = rj.c;

  // ---- rare: exact sums of the ambiguous positions, then CTA 0 of the group decides ----
  if (seq_tasks > 0) {
    for (int i = 0; i < g; ++i) {  // block-uniform
      if ((__ldcg((const unsigned char*)&ws.status[(long long)b * g + i]) & 3) != ST_AMBIG) continue;
      const long long r1 = (long long)b * rps + i, r2 = (long long)b * rps + rj.nT + i;
      const void* prow = seq_row_ptr<DT>(rj, b, i);
      const void* qrow = seq_row_ptr<DT>(rj, b, rj.nT + i);
      u64 sp = 0, sq = 0;
      pair_sums<DT, false>(prow, qrow, (((size_t)prow) & 15) == 0, (((size_t)qrow) & 15) == 0, V, c,
                           ldcg_rowout(&rj.out[r1]).mc, ldcg_rowout(&rj.out[r2]).mc, s0, s1, nullptr, sp, sq);
      block_sum2_u64(sp, sq, sh.sh64);
      if (tid == 0) {
        if (sp) atomicAdd(&ws.acc[r1], sp);
        if (sq) atomicAdd(&ws.acc[r2], sq);
      }
    }
    if (tid == 0) {
      __threadfence();
      atomicAdd(&ws.exact_done[b], 1);
      sh.sh_ok = 1;
      if (ch == 0) sh.sh_ok = spin_until(&ws.exact_done[b], TF_CH, ws.abort) ? 1 : 0;
    }
    cta_sync();
    if (!sh.sh_ok) return false;
    if (ch == 0 && tid < 32) {
      decide_sequence<DT>(job, ws, b);
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicExch(&ws.decided[b], 1);
      }
    }
    if (tid == 0) sh.sh_ok = spin_until(&ws.decided[b], 1, ws.abort) ? 1 : 0;
    cta_sync();
    if (!sh.sh_ok) return false;
  }
  if (!have_rec || seq_tasks > 0) {  // (the record handed in predates the exact decisions of an ambiguous sequence)
    rec = __ldcg((const int4*)(ws.samp + b * SAMP_N));
    mcq_bits = __ldcg(&ws.samp[b * SAMP_N + 4]);
  }
  const int mode = rec.y, prow_i = rec.z;
  const float mcp = __int_as_float(rec.w), mcq = __int_as_float(mcq_bits);
  if (mode == 0) return true;

  const long long r1 = (long long)b * rps + prow_i;
  const void* prowp = seq_row_ptr<DT>(rj, b, prow_i);
  const bool pal = (((size_t)prowp) & 15) == 0;
  u64* part = ws.part + (size_t)b * ws.nseg_pad;
  u64 tot = 0;
  float best = (mode == 2) ? 0.0f : -1.0f;
  int bidx = 0x7FFFFFFF;
  if (mode == 1) {
    // bonus / target row itself: the sampling weights are the canonical weights, one evaluation suffices
    const RowOut rp = resolved_row(rj, ws, r1);
    partial_loop<DT, false, GREEDY, false>(prowp, prowp, rp, rp, pal, pal, V, c, s0, s1, part, tot, best, bidx);
  } else {
    const long long r2 = (long long)b * rps + rj.nT + prow_i;
    const void* qrowp = seq_row_ptr<DT>(rj, b, rj.nT + prow_i);
    const bool qal = (((size_t)qrowp) & 15) == 0;
    // ---- phase A: canonical weights of my slice -> shared memory; partial normalisers -> group ----
    u64 sp = 0, sq = 0;
    pair_sums<DT, true>(prowp, qrowp, pal, qal, V, c, mcp, mcq, s0, s1, ecache, sp, sq);
    block_sum2_u64(sp, sq, sh.sh64);
    if (tid == 0) {
      if (sp) atomicAdd(&ws.acc2[2 * b], sp);
      if (sq) atomicAdd(&ws.acc2[2 * b + 1], sq);
      __threadfence();
      atomicAdd(&ws.fin_done[b], 1);
      sh.sh_ok = spin_until(&ws.fin_done[b], TF_CH, ws.abort) ? 1 : 0;
      dbg_stamp_max(ws, 16 + b * 8 + 4);
      const u64 Sp = __ldcg(&ws.acc2[2 * b]), Sq = __ldcg(&ws.acc2[2 * b + 1]);
      sh.sh_S[0] = Sp; sh.sh_S[1] = Sq;
      // resolved_row() reads the normalisers from acc[]: every CTA of the group stores the same values
      ws.acc[r1] = Sp;
      ws.acc[r2] = Sq;
    }
    cta_sync();
    if (!sh.sh_ok) return false;
    const float invp = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(sh.sh_S[0]), 0x1p-40f));
    const float invq = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(sh.sh_S[1]), 0x1p-40f));
    if (ch == 0 && tid == 0 && prow_i < g) {
      // the deciding position reports its exact probabilities (as the exact_rows pipeline does)
      RowOut rp = ldcg_rowout(&rj.out[r1]), rq = ldcg_rowout(&rj.out[r2]);
      rp.inv = invp; rq.inv = invq;
      const long long* toks = job.draft_tokens + (long long)b * g;
      const int tok = (int)min(max(toks[prow_i], 0ll), (long long)V - 1);
      job.p_tok[(long long)b * g + prow_i] = row_prob<DT>(rp, prowp, tok, c);
      job.q_tok[(long long)b * g + prow_i] = row_prob<DT>(rq, qrowp, tok, c);
    }
    // ---- phase B: residual partial sums from the cached weights ----
    const float ip20 = __fmul_rn(invp, 1048576.0f), iq20 = __fmul_rn(invq, 1048576.0f);  // (exact)
    const float2 ip2 = make_float2(ip20, ip20), iq2 = make_float2(iq20, iq20);
    if (w < TF_T / 32)
    for (int seg = s0 + w; seg < s1; seg += TF_T / 32) {
      const float4* src = ecache + (size_t)(seg - s0) * 128 + lane;
      const int j0 = (seg * 32 + lane) * 8;
      u64 s = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 rr = resid2_s60(src[k * 32], ip2, iq2);  // max(0, P - Q) * 2^60 of tokens j0 + 2k, j0 + 2k + 1
        const float v0 = rr.x, v1 = rr.y;
        s += __float2ull_rz(v0) + __float2ull_rz(v1);           // == fix60(max(0, P - Q))
        if (GREEDY) {
          if (j0 + 2 * k < V && v0 > best) { best = v0; bidx = j0 + 2 * k; }
          if (j0 + 2 * k + 1 < V && v1 > best) { best = v1; bidx = j0 + 2 * k + 1; }
        }
      }
      s = warp_sum_u64(s);
      if (lane == 0) { part[seg] = s; tot += s; }
    }
  }
  // ---- tail (as sample_partial_kernel): totals, then the last CTA of the group finalizes ----
  tot = block_sum_u64(tot, sh.sh64);
  if (GREEDY) {
    u64 key = (bidx == 0x7FFFFFFF) ? 0ull : (((u64)__float_as_uint(best)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)bidx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, key, o); key = t > key ? t : key; }
    if (lane == 0 && key) atomicMax(&ws.best[b], key);
  }
  cta_sync();
  if (tid == 0) {
    if (tot) atomicAdd(&ws.tot[b], tot);
    __threadfence();
    sh.sh_last = (atomicAdd(&ws.part_done[b], 1) == TF_CH - 1) ? 1 : 0;
  }
  cta_sync();
  if (sh.sh_last) {
    __threadfence();
    if (tid == 0) dbg_stamp_max(ws, 16 + b * 8 + 5);
    finalize_sequence<DT>(job, ws, b, sh.sh64, sh.shf, sh.shi, &sh.s_res);
    if (tid == 0) dbg_stamp_max(ws, 16 + b * 8 + 6);
  }
  return true;
}

template <int DT, bool GREEDY>
__global__ void __launch_bounds__(TF_T, 4) tail_fused_kernel(DecideJob job, HybridWs ws, int segs_per_cta, int TF_CH) {
  extern __shared__ __align__(16) float4 ecache[];
  __shared__ TailSh sh;
  __shared__ int sh_ticket;
  // (the ticket counter was zeroed by the memset at the head of the call, long before the plan kernel: tickets can
  // be taken while the plan kernel is still running; its results are only read after the dependency wait)
  if (threadIdx.x == 0) sh_ticket = atomicAdd(ws.ticket, 1);
  grid_dependency_wait();
  __syncthreads();
  const int b = sh_ticket / TF_CH, ch = sh_ticket - b * TF_CH;
  tail_item<DT, GREEDY>(job, ws, b, ch, TF_CH, segs_per_cta, ecache, sh, ws.seq_tasks[b], false, make_int4(0, 0, 0, 0), 0);
}
