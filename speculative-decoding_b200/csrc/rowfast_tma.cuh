// rowfast_tma.cuh -- the HBM-bound row-statistics kernel as a TMA bulk-copy pipeline (included inside
// namespace specdec by hybrid.cuh).
//
// Persistent CTAs (4 per SM): one producer warp streams each logit row HBM -> shared memory with
// cp.async.bulk (1-D TMA, SASS UBLKCP) into a 3-stage x 16 KB ring per CTA (4 CTAs / SM), completion signalled through
// mbarrier transaction counts; 8 consumer warps read the stages with conflict-free 16-byte LDS (4 vectors per
// thread and stage: measured 0.1044 ms vs 0.1095 ms with 6 x 8 KB stages -- fewer barrier round trips) and
// keep the online (max, sum of MUFU ex2) per thread.  ~190 KB of loads are in flight per SM
// independent of what the consumers are doing (block reductions, row epilogues), which is what the
// plain LDG version (rowfast_kernel) could not sustain.  Used when every row is 16-byte aligned and a
// multiple of 16 bytes; otherwise rowfast_kernel runs.
#pragma once

constexpr int TS_CONSUMERS = 256;
constexpr int TS_THREADS = TS_CONSUMERS + 32;
#ifndef TS_STAGES_V
#define TS_STAGES_V 3
#endif
#ifndef TS_STAGE_BYTES_V
#define TS_STAGE_BYTES_V 16384
#endif
constexpr int TS_STAGES = TS_STAGES_V;
constexpr int TS_STAGE_BYTES = TS_STAGE_BYTES_V;
constexpr int TS_SMEM = TS_STAGES * TS_STAGE_BYTES;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// streaming read: the logits are consumed once by this kernel, so they are marked evict-first in L2
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar,
                                             unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// online (max, sum) update with one 16-byte vector
template <int DT>
__device__ __forceinline__ void online16(const uint4 raw, float& m, float& s, const float c) {
  if (DT == DT_F32) {
    const float x0 = __uint_as_float(raw.x), x1 = __uint_as_float(raw.y), x2 = __uint_as_float(raw.z),
                x3 = __uint_as_float(raw.w);
    const float vm = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
    if (vm > m) { s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, vm), c))); m = vm; }
    const float mc = (m > -INFINITY) ? __fmul_rn(m, c) : 0.0f;  // (only -inf so far: -inf*c + inf would be NaN)
    s = __fadd_rn(s, ex2_approx(__fmaf_rn(x0, c, -mc)));
    s = __fadd_rn(s, ex2_approx(__fmaf_rn(x1, c, -mc)));
    s = __fadd_rn(s, ex2_approx(__fmaf_rn(x2, c, -mc)));
    s = __fadd_rn(s, ex2_approx(__fmaf_rn(x3, c, -mc)));
  } else {
    float x[8];
    const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (DT == DT_BF16) {
        x[2 * k] = __uint_as_float(w[k] << 16);
        x[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        x[2 * k] = f.x; x[2 * k + 1] = f.y;
      }
    }
    const float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
    if (vm > m) { s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, vm), c))); m = vm; }
    // packed fp32x2 (FFMA2 / FADD2): half the issue slots for the exponent arguments and the accumulation
    const float nmc_ = (m > -INFINITY) ? -__fmul_rn(m, c) : 0.0f;
    const float2 c2 = make_float2(c, c), nmc2 = make_float2(nmc_, nmc_);
    float2 acc = make_float2(s, 0.0f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = __ffma2_rn(make_float2(x[2 * k], x[2 * k + 1]), c2, nmc2);
      acc = __fadd2_rn(acc, make_float2(ex2_approx(t.x), ex2_approx(t.y)));
    }
    s = __fadd_rn(acc.x, acc.y);
  }
}

// MODE 1 (NUC): pre-pass of the top-p rows (launch_rowstats): the caller passes c = log2(e) (T = 1 masses) and the row
// epilogue additionally emits what nucleus_fast_kernel needs to go straight to its candidate sweep -- RowOut.inv =
// the MUFU mass S (not 1/S), RowOut.cut = candidate threshold (min over the 4-lane-group maxima of the 256 consumer
// threads: >= 64 elements lie above it), RowOut.Sfix = bits of the mass carried by the 256 per-thread maxima.
// Maximum of the 8 packed 16-bit elements of one vector, kept PACKED (two lanes): HMNMX2 on bf16x2 / f16x2.
template <int DT>
__device__ __forceinline__ unsigned packed_max4(const uint4 raw, unsigned pm) {
  if (DT == DT_BF16) {
    __nv_bfloat162 a = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x), *reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    __nv_bfloat162 b = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&raw.z), *reinterpret_cast<const __nv_bfloat162*>(&raw.w));
    a = __hmax2(__hmax2(a, b), *reinterpret_cast<const __nv_bfloat162*>(&pm));
    return *reinterpret_cast<unsigned*>(&a);
  } else {
    __half2 a = __hmax2(*reinterpret_cast<const __half2*>(&raw.x), *reinterpret_cast<const __half2*>(&raw.y));
    __half2 b = __hmax2(*reinterpret_cast<const __half2*>(&raw.z), *reinterpret_cast<const __half2*>(&raw.w));
    a = __hmax2(__hmax2(a, b), *reinterpret_cast<const __half2*>(&pm));
    return *reinterpret_cast<unsigned*>(&a);
  }
}
template <int DT>
__device__ __forceinline__ float packed_max_to_float(unsigned pm) {
  if (DT == DT_BF16) return fmaxf(__uint_as_float(pm << 16), __uint_as_float(pm & 0xFFFF0000u));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pm));
  return fmaxf(f.x, f.y);
}
// sum of the 8 MUFU weights of one 16-bit vector under a maximum m that already covers it (no max logic)
template <int DT>
__device__ __forceinline__ void accumulate16(const uint4 raw, const float2 c2, const float2 nmc2, float2& acc) {
  const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float2 x;
    if (DT == DT_BF16) { x.x = __uint_as_float(w[k] << 16); x.y = __uint_as_float(w[k] & 0xFFFF0000u); }
    else x = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
    const float2 t = __ffma2_rn(x, c2, nmc2);
    acc = __fadd2_rn(acc, make_float2(ex2_approx(t.x), ex2_approx(t.y)));
  }
}

// MODE 2 (ARGMAX, 16-bit rows): greedy n-gram verify (rowfast_argmax_kernel's job at TMA speed): additionally the first
// index of the row maximum.  RowOut.Sfix = index | ambiguous << 32; "ambiguous" (the runner-up could be within 4e-6
// exponent units of the maximum, where the canonical arg-max over the weights may differ from the arg-max over the
// logits) is decided from the spacing of the 16-bit format at the maximum instead of tracking the runner-up.
template <int DT, int MODE = 0>
__global__ void __launch_bounds__(TS_THREADS, 4) rowfast_tma_kernel(DecideJob dj, HybridWs ws) {
  constexpr bool NUC = (MODE == 1), AMAX = (MODE == 2);
  const RowJob& job = dj.rj;
  // PDL: let the next kernel of the stream (the row kernel of the next chunk, which never waits, or plan_kernel, which
  // does) be scheduled as soon as SM resources free up instead of after this grid's last CTA
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(8) unsigned long long full_bar[TS_STAGES], empty_bar[TS_STAGES];
  __shared__ float sh_m[2][8], sh_s[2][8], sh_q[2][8], sh_w[2][8];
  __shared__ unsigned sh_i[2][8];
  // Rows are CLAIMED from a counter (ws.r_claim, zeroed at the head of the call) instead of being assigned by
  // blockIdx: a CTA that is scheduled late -- another kernel holds a slot of its SM (the NCCL all-gather of the
  // previous step at 8 GPUs: 176 -> 244 us per step with static rows), or the grid does not divide the rows
  // (1152 rows on 444 CTAs) -- finds less work left instead of finishing its fixed share alone after everyone else.
  // The producer publishes each claimed row id in row_id[row number mod 8] before the row's first bulk copy (visible
  // to the consumers through the stage's mbarrier); -1 ends the CTA.  The producer is at most TS_STAGES stages, hence
  // at most TS_STAGES rows (one-stage rows of a tiny vocabulary), ahead of the slowest consumer warp: 8 slots suffice.
  // Without a counter (ws.r_claim == nullptr): static rows.
  static_assert(TS_STAGES + 2 <= 8, "row_id ring too small");
  __shared__ long long row_id[8];
  int* const r_claim = ws.r_claim;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned row_bytes = (unsigned)job.V * ((DT == DT_F32) ? 4u : 2u);
  const int nst = (int)((row_bytes + TS_STAGE_BYTES - 1) / TS_STAGE_BYTES);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < TS_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], TS_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == TS_CONSUMERS / 32) {
    // ---------------- producer warp: one elected lane issues the bulk copies ----------------
    if (lane == 0) {
      int stage = 0;
      unsigned phase = 0;
      const unsigned long long policy = l2_evict_first_policy();
      int rpar = 0;
      long long r = r_claim ? (long long)atomicAdd(r_claim, 1) : (long long)blockIdx.x;
      for (; r < job.R; rpar = (rpar + 1) & 7) {
        // the next claim is issued before this row is streamed: its round trip hides behind 17 bulk copies
        const long long r_next = r_claim ? (long long)atomicAdd(r_claim, 1) : r + gridDim.x;
        const char* base = (const char*)row_ptr<DT>(job, r);
        for (int k = 0; k < nst; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (k == 0) row_id[rpar] = r;  // (released to the consumers by the arrive below)
          const unsigned off = (unsigned)k * TS_STAGE_BYTES;
          const unsigned nb = min((unsigned)TS_STAGE_BYTES, row_bytes - off);
          mbar_expect_tx(&full_bar[stage], nb);
          tma_bulk_g2s(ring + stage * TS_STAGE_BYTES, base + off, nb, &full_bar[stage], policy);
          if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
        }
        r = r_next;
      }
      // no row left: complete one more phase without data so that the consumers wake up and read the end marker
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      row_id[rpar] = -1;
      mbar_arrive(&full_bar[stage]);
    }
    return;
  }
  // ---------------- consumer warps ----------------
  const float c = job.c;
  int stage = 0, par = 0;
  unsigned phase = 0;
  for (int rseq = 0;; rseq = (rseq + 1) & 7, par ^= 1) {
    mbar_wait(&full_bar[stage], phase);  // first stage of the next row, or the end marker
    const long long r = *(volatile long long*)&row_id[rseq];
    if (r < 0) break;
    float m = -INFINITY, s = 0.0f;
    unsigned first = 0xFFFFFFFFu;  // AMAX: first index of this thread's maximum
    for (int k = 0; k < nst; ++k) {
      if (k > 0) mbar_wait(&full_bar[stage], phase);
      const unsigned off = (unsigned)k * TS_STAGE_BYTES;
      const int nvec = (int)(min((unsigned)TS_STAGE_BYTES, row_bytes - off) >> 4);
      const uint4* sp = reinterpret_cast<const uint4*>(ring + stage * TS_STAGE_BYTES);
      constexpr int VPT = TS_STAGE_BYTES / 16 / TS_CONSUMERS;  // 16-byte vectors per consumer thread and stage
      uint4 a[VPT];
#pragma unroll
      for (int q = 0; q < VPT; ++q)
        if (tid + q * TS_CONSUMERS < nvec) a[q] = sp[tid + q * TS_CONSUMERS];
      if (DT == DT_F32) {
#pragma unroll
        for (int q = 0; q < VPT; ++q)
          if (tid + q * TS_CONSUMERS < nvec) online16<DT>(a[q], m, s, c);
      } else {
        // 16-bit rows: the maximum of the thread's (up to) VPT vectors of the stage first, on the packed words
        // (HMNMX2, 4 per vector), ONE rescale test per stage, then the weights without any max logic
        unsigned pm = (DT == DT_BF16) ? 0xFF80FF80u : 0xFC00FC00u;  // (-inf, -inf)
#pragma unroll
        for (int q = 0; q < VPT; ++q)
          if (tid + q * TS_CONSUMERS < nvec) pm = packed_max4<DT>(a[q], pm);
        const float vm = packed_max_to_float<DT>(pm);
        if (vm > m) {
          s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, vm), c)));
          m = vm;
          if (AMAX) {  // rare: locate the first element of the stage that equals the new maximum
            first = 0xFFFFFFFFu;
#pragma unroll
            for (int q = VPT - 1; q >= 0; --q)
              if (tid + q * TS_CONSUMERS < nvec) {
                const unsigned w[4] = {a[q].x, a[q].y, a[q].z, a[q].w};
                const unsigned j0 = ((off >> 4) + (unsigned)(tid + q * TS_CONSUMERS)) * 8u;
#pragma unroll
                for (int e = 7; e >= 0; --e) {
                  const unsigned h = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xFFFFu);
                  const float x = (DT == DT_BF16) ? __uint_as_float(h << 16) : __half2float(__ushort_as_half((unsigned short)h));
                  if (x == vm) first = j0 + (unsigned)e;
                }
              }
          }
        }
        const float nmc = (m > -INFINITY) ? -__fmul_rn(m, c) : 0.0f;  // (a thread that has only seen -inf: no NaN)
        const float2 c2 = make_float2(c, c), nmc2 = make_float2(nmc, nmc);
        float2 acc = make_float2(s, 0.0f);
#pragma unroll
        for (int q = 0; q < VPT; ++q)
          if (tid + q * TS_CONSUMERS < nvec) accumulate16<DT>(a[q], c2, nmc2, acc);
        s = __fadd_rn(acc.x, acc.y);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
    }
    // row epilogue among the 256 consumer threads (named barrier 1; the producer keeps prefetching).
    // Scratch is double-buffered by row parity, so one barrier per row suffices.
    float wm = warp_max_f(m);
    const float resc = (m > -INFINITY) ? ex2_approx(__fmul_rn(__fsub_rn(m, wm), c)) : 0.0f;  // weight of my maximum
    s = (m > -INFINITY) ? __fmul_rn(s, resc) : 0.0f;  // (a thread that saw only -inf carries NaN in s: dropped)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { sh_m[par][warp] = wm; sh_s[par][warp] = s; }
    if (AMAX) {
      const unsigned wf = __reduce_min_sync(0xffffffffu, (m == wm) ? first : 0xFFFFFFFFu);
      if (lane == 0) sh_i[par][warp] = wf;
    }
    if (NUC) {
      float qm = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));  // 4-lane groups: 64 groups of ~V/64 elements,
      qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, 2));      // the same statistics as nucleus_fast_kernel's own sweep
      const float wq = warp_min_f(qm);
      float wt = resc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wt += __shfl_xor_sync(0xffffffffu, wt, o);
      if (lane == 0) { sh_q[par][warp] = wq; sh_w[par][warp] = wt; }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(TS_CONSUMERS) : "memory");
    if (warp == 0) {
      if (lane == 0) {
        float M = sh_m[par][0];
#pragma unroll
        for (int w = 1; w < TS_CONSUMERS / 32; ++w) M = fmaxf(M, sh_m[par][w]);
        float S = 0.0f;
#pragma unroll
        for (int w = 0; w < TS_CONSUMERS / 32; ++w)
          S += (sh_m[par][w] > -INFINITY) ? __fmul_rn(sh_s[par][w], ex2_approx(__fmul_rn(__fsub_rn(sh_m[par][w], M), c))) : 0.0f;
        RowOut o;
        o.m = M; o.mc = __fmul_rn(M, c); o.inv = __fdiv_rn(1.0f, S);
        o.cut = -INFINITY; o.jcut = job.V; o.flags = 0; o.Sfix = 0;
        if (NUC) {
          float th = sh_q[par][0], Wt = 0.0f;
#pragma unroll
          for (int w = 0; w < TS_CONSUMERS / 32; ++w) {
            th = fminf(th, sh_q[par][w]);
            Wt += (sh_m[par][w] > -INFINITY) ? __fmul_rn(sh_w[par][w], ex2_approx(__fmul_rn(__fsub_rn(sh_m[par][w], M), c))) : 0.0f;
          }
          o.inv = S; o.cut = th; o.Sfix = (u64)__float_as_uint(Wt);
        }
        if (AMAX) {
          unsigned fi = 0xFFFFFFFFu;
#pragma unroll
          for (int w = 0; w < TS_CONSUMERS / 32; ++w)
            if (sh_m[par][w] == M) fi = min(fi, sh_i[par][w]);
          // smallest possible gap below M in this 16-bit format: half an ulp of M (M a power of two)
          const int ex = (int)((__float_as_uint(M) >> 23) & 0xFFu) - ((DT == DT_BF16) ? 8 : 11);
          const float gap = (ex > 0 && ex < 255) ? __uint_as_float((unsigned)ex << 23) : 0.0f;
          const bool amb = !(M > -INFINITY) || !(M < INFINITY) || fi >= (unsigned)job.V || !(__fmul_rn(gap, c) >= 4e-6f);
          o.Sfix = (u64)fi | (amb ? (1ull << 32) : 0ull);
        }
        job.out[r] = o;
      }
    }
  }
}
