// mega.cuh -- the verify step of the plain sampling modes as ONE persistent, cooperative launch (included by
// hybrid.cuh inside namespace specdec).
//
// The three-launch pipeline (row kernel -> plan -> fused tail, hybrid.cuh) leaves the issue-bound exact tail exposed
// behind the HBM-bound row pass and re-reads the deciding row pair from HBM.  Here both run side by side on every SM:
//
//   R CTAs (blockIdx < n_r, 2 per SM): the TMA row pipeline of rowfast_tma.cuh over row SLICES ("units" of
//     unit_stages x 16 KB) in sequence-major order, so all R CTAs of the GPU work on the same handful of sequences and
//     a sequence's rows complete within microseconds of each other.  Each unit publishes (max, MUFU sum) of its slice
//     and bumps the sequence's unit counter (red.release).
//   X CTAs (2 per SM): warps 0-7 = compute group, warp 8 = service warp.
//     service warp: (a) PLANS the next unplanned sequence whose units are complete -- merges the slice statistics into
//       RowOut records, runs plan_sequence (accept tests with the 1e-3 margin, hybrid.cuh) and publishes the plan
//       (st.release plan_done[b]); claimed with a CAS, never blocks.  (b) SCOUTS for its own compute group: claims the
//       next exact item (sequence b, slice s) in order, waits until b is planned, and hands the plan record over
//       through a two-slot shared-memory mailbox, so the compute group never sees a global-memory round trip for it.
//     compute group: tail_item() of tail_fused.cuh per item -- canonical weights of its slice of the deciding row pair
//       (read back through L2: the pair was streamed microseconds ago with the normal eviction policy) cached in the
//       48 KB that are the TMA ring in an R CTA, partial normalisers to the group by u64 atomics, residual partial sums
//       from the cache, last slice of a sequence finalizes (scan + token).
//   R CTAs whose units are done turn into X CTAs (their producer warp becomes the service warp), so the drain of the
//   last sequences runs on every warp of the GPU.
//
// Forward progress: the launch is cooperative (all CTAs co-resident, the launch fails otherwise).  R never waits for
// anything but its own TMA.  Plans wait for R only.  Items are claimed in order and a compute group holds at most two;
// its waits (siblings' partial sums of the same sequence) only involve items that precede its next claim, so with
// more than 2 S compute groups the lowest waiting sequence always completes.  Every wait is bounded (spin_until).
// Results are bit-identical to the three-launch pipeline: the same integers are summed.
#pragma once

struct MegaCfg {
  int n_r;          // CTAs with the R role (row-slice lanes)
  int n_x;          // CTAs that start in the X role
  int r_per_sm;     // R CTAs per SM
  int U;            // units (slices) per logit row, <= MG_MAXU
  int unit_stages;  // TMA stages per unit
  int S;            // exact items (slices of the deciding row pair) per sequence
  int spc;          // 256-element segments per exact item
  int B;            // sequences
  int keep_l2;      // stream the rows with the normal L2 eviction policy (re-read of the deciding pair hits L2)
};

__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_normal_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

#ifndef SPECDEC_POLL_NS
#define SPECDEC_POLL_NS 40  // pause between polls of the exchange words (tuning builds: -DSPECDEC_POLL_NS=...)
#endif
#ifndef MG_NS_V
#define MG_NS_V 2
#endif
constexpr int MG_NS = MG_NS_V;  // 256-pair segments in flight per warp in the exact items (register budget: 56)

struct MegaSh {
  TailSh tail;
  volatile int credit;         // compute group -> service warp: "my current item waits for nobody any more, claim the next"
  volatile int slot_state[2];  // mailbox service warp -> compute group: 0 empty, 1 filled
  int slot_item[2];            // item index, < 0: no more items (-2: aborted)
  int slot_rec[2][8];          // the sequence's plan record (HybridWs::samp)
};

// ---------------------------------------------------------------------------------------------
// R role: TMA row-slice streaming (the consumer arithmetic is rowfast_tma_kernel's)
// ---------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ void mega_stream_rows(const DecideJob& dj, const HybridWs& ws, const MegaCfg& cfg, const int r_lane,
                                                 unsigned char* ring, unsigned long long* full_bar,
                                                 unsigned long long* empty_bar, float (*sh_m)[8], float (*sh_s)[8]) {
  const RowJob& job = dj.rj;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned row_bytes = (unsigned)job.V * ((DT == DT_F32) ? 4u : 2u);
  const unsigned unit_bytes = (unsigned)cfg.unit_stages * TS_STAGE_BYTES;
  const int U = cfg.U;
  const long long NU = job.R * U;
  if (warp == TS_CONSUMERS / 32) {
    // ---------------- producer warp: one elected lane issues the bulk copies ----------------
    if (lane == 0) {
      int stage = 0;
      unsigned phase = 0;
      const unsigned long long policy = cfg.keep_l2 ? l2_evict_normal_policy() : l2_evict_first_policy();
      for (long long u = r_lane; u < NU; u += cfg.n_r) {
        const long long r = u / U;
        const unsigned beg = (unsigned)(u - r * U) * unit_bytes, len = min(unit_bytes, row_bytes - beg);
        const char* base = (const char*)row_ptr<DT>(job, r) + beg;
        for (unsigned off = 0; off < len; off += TS_STAGE_BYTES) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const unsigned nb = min((unsigned)TS_STAGE_BYTES, len - off);
          mbar_expect_tx(&full_bar[stage], nb);
          tma_bulk_g2s(ring + stage * TS_STAGE_BYTES, base + off, nb, &full_bar[stage], policy);
          if (++stage == TS_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
    return;
  }
  // ---------------- consumer warps ----------------
  const float c = job.c;
  int stage = 0, par = 0;
  unsigned phase = 0;
  for (long long u = r_lane; u < NU; u += cfg.n_r, par ^= 1) {
    const long long r = u / U;
    const int sl = (int)(u - r * U);
    const unsigned beg = (unsigned)sl * unit_bytes, len = min(unit_bytes, row_bytes - beg);
    float m = -INFINITY, s = 0.0f;
    for (unsigned off = 0; off < len; off += TS_STAGE_BYTES) {
      mbar_wait(&full_bar[stage], phase);
      const int nvec = (int)(min((unsigned)TS_STAGE_BYTES, len - off) >> 4);
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
        // maximum of the thread's vectors of the stage on the packed words (HMNMX2), ONE rescale test per stage
        unsigned pm = (DT == DT_BF16) ? 0xFF80FF80u : 0xFC00FC00u;  // (-inf, -inf)
#pragma unroll
        for (int q = 0; q < VPT; ++q)
          if (tid + q * TS_CONSUMERS < nvec) pm = packed_max4<DT>(a[q], pm);
        const float vm = packed_max_to_float<DT>(pm);
        if (vm > m) {
          s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, vm), c)));
          m = vm;
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
    // unit epilogue among the 256 consumer threads (named barrier 1; the producer keeps prefetching);
    // scratch double-buffered by unit parity, so one barrier per unit suffices
    const float wm = warp_max_f(m);
    const float resc = (m > -INFINITY) ? ex2_approx(__fmul_rn(__fsub_rn(m, wm), c)) : 0.0f;
    s = (m > -INFINITY) ? __fmul_rn(s, resc) : 0.0f;  // (a thread that saw only -inf carries NaN in s: dropped)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { sh_m[par][warp] = wm; sh_s[par][warp] = s; }
    asm volatile("bar.sync 1, %0;" ::"n"(TS_CONSUMERS) : "memory");
    if (tid == 0) {
      float M = sh_m[par][0];
#pragma unroll
      for (int w = 1; w < TS_CONSUMERS / 32; ++w) M = fmaxf(M, sh_m[par][w]);
      float S = 0.0f;
#pragma unroll
      for (int w = 0; w < TS_CONSUMERS / 32; ++w)
        S += (sh_m[par][w] > -INFINITY) ? __fmul_rn(sh_s[par][w], ex2_approx(__fmul_rn(__fsub_rn(sh_m[par][w], M), c))) : 0.0f;
      // one 8-byte store, valid by itself: the records are zeroed per call and (0, 0) is not a possible value
      // (S >= 1 unless the whole slice is -inf, which gives (-inf, 0)) -- no counter, no fence, nothing to wait for
      st_relaxed_u64(reinterpret_cast<u64*>(ws.rpart + r * MG_MAXU + sl),
                     ((u64)__float_as_uint(S) << 32) | (u64)__float_as_uint(M));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plan of sequence b (one warp): slice statistics -> RowOut records -> plan_sequence -> publish
// ---------------------------------------------------------------------------------------------
// true iff every slice record of sequence b has been published (one warp, non-blocking)
__device__ __forceinline__ bool mega_rows_ready(const HybridWs& ws, const MegaCfg& cfg, int b, int rps) {
  const int lane = threadIdx.x & 31;
  const u64* base = reinterpret_cast<const u64*>(ws.rpart) + (size_t)b * rps * MG_MAXU;
  bool ok = true;
  for (int i0 = 0; i0 < rps * cfg.U; i0 += 32) {
    const int i = i0 + lane;
    if (i < rps * cfg.U) ok = ok && (ld_relaxed_u64(base + (size_t)(i / cfg.U) * MG_MAXU + (i % cfg.U)) != 0ull);
  }
  return __all_sync(0xffffffffu, ok);
}

template <int DT>
__device__ __forceinline__ void mega_plan(const DecideJob& job, const HybridWs& ws, const MegaCfg& cfg, int b) {
  const RowJob& rj = job.rj;
  const int lane = threadIdx.x & 31, rps = rj.nT + rj.nD, U = cfg.U;
  const float c = rj.c;
  if (lane == 0) dbg_stamp_max(ws, 16 + b * 8 + 0);
  for (int k = lane; k < rps; k += 32) {
    const long long r = (long long)b * rps + k;
    const u64* pp = reinterpret_cast<const u64*>(ws.rpart + r * MG_MAXU);
    float vm[MG_MAXU], vs[MG_MAXU];
    float M = -INFINITY;
#pragma unroll
    for (int sl = 0; sl < MG_MAXU; ++sl) {
      const u64 w = (sl < U) ? ld_relaxed_u64(pp + sl) : 0ull;
      vm[sl] = (sl < U) ? __uint_as_float((unsigned)w) : -INFINITY;
      vs[sl] = __uint_as_float((unsigned)(w >> 32));
      M = fmaxf(M, vm[sl]);
    }
    float S = 0.0f;
#pragma unroll
    for (int sl = 0; sl < MG_MAXU; ++sl)
      S += (vm[sl] > -INFINITY) ? __fmul_rn(vs[sl], ex2_approx(__fmul_rn(__fsub_rn(vm[sl], M), c))) : 0.0f;
    RowOut o;
    o.m = M; o.mc = __fmul_rn(M, c); o.inv = __fdiv_rn(1.0f, S);
    o.cut = -INFINITY; o.jcut = rj.V; o.flags = 0; o.Sfix = 0;
    rj.out[r] = o;
  }
  __syncwarp();
  plan_sequence<DT>(job, ws, b);
  __syncwarp();
  __threadfence();
  if (lane == 0) { st_release_gpu(&ws.plan_done[b], 1); dbg_stamp_max(ws, 16 + b * 8 + 1); }
}

// ---------------------------------------------------------------------------------------------
// X role, service warp: plans + mailbox
// ---------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ void mega_service_warp(const DecideJob& job, const HybridWs& ws, const MegaCfg& cfg, MegaSh& sh,
                                                  const int x_rank) {
  const int lane = threadIdx.x & 31;
  const int rps = job.rj.nT + job.rj.nD;
  const int total_items = cfg.B * cfg.S;
  int slot = 0, claimed = -1;
  // (R CTAs that turned into X CTAs after their row streaming do not plan: every sequence is planned long before)
  int my_plan = (x_rank >= 0) ? x_rank : cfg.B;
  bool claims_done = false, plans_done = false;
  unsigned idle = 0;
  while (!(claims_done && plans_done)) {
    bool progress = false;
    if (!plans_done) {  // (a) plan the next unplanned sequence if its row statistics are complete
      // sequences are planned by the service warps of the X CTAs in a fixed round-robin (my_plan, my_plan + NX, ...): a
      // warp polls only the records of its own next sequence, nothing is serialised through a shared counter
      if (my_plan >= cfg.B) plans_done = true;
      else if (mega_rows_ready(ws, cfg, my_plan, rps)) {
        mega_plan<DT>(job, ws, cfg, my_plan);
        my_plan += cfg.n_x;
        progress = true;
      }
    }
    if (!claims_done) {  // (b) next item for my compute group
      int st = 0;  // 1: delivered an item, 2: delivered the end marker
      if (lane == 0) {
        // An item is claimed only once the compute group's current item is past its last wait for sibling slices: a
        // CTA never sits on an unstarted item while blocked, so two items of one sequence in one CTA cannot deadlock.
        if (claimed < 0 && sh.credit != 0 && sh.slot_state[slot] == 0) { sh.credit = 0; claimed = atomicAdd(ws.x_next, 1); }
        if (claimed >= total_items) {
          sh.slot_item[slot] = -1;
          __threadfence_block();
          sh.slot_state[slot] = 1;
          st = 2;
        } else if (claimed >= 0) {
          const int b = claimed / cfg.S;
          if (ld_acquire_gpu(&ws.plan_done[b]) != 0) {
            const int4 r0 = __ldcg((const int4*)(ws.samp + b * SAMP_N)), r1 = __ldcg((const int4*)(ws.samp + b * SAMP_N) + 1);
            int* d = sh.slot_rec[slot];
            d[0] = r0.x; d[1] = r0.y; d[2] = r0.z; d[3] = r0.w; d[4] = r1.x; d[5] = r1.y;
            sh.slot_item[slot] = claimed;
            __threadfence_block();
            sh.slot_state[slot] = 1;
            st = 1;
          }
        }
      }
      st = __shfl_sync(0xffffffffu, st, 0);
      if (st == 2) claims_done = true;
      if (st) { slot ^= 1; claimed = -1; progress = true; }
      claimed = __shfl_sync(0xffffffffu, claimed, 0);
    }
    if (progress) { idle = 0; continue; }
    if ((++idle & 255u) == 0u) {  // bounded: an aborted launch (or > ~1 s without progress) ends the service warp
      int ab = 0;
      if (lane == 0) {
        ab = ld_acquire_gpu(ws.abort);
        if (!ab && idle > (1u << 22)) { atomicExch(ws.abort, 1); ab = 1; }
      }
      if (__shfl_sync(0xffffffffu, ab, 0)) return;  // (the compute group watches the abort word itself)
    }
    __nanosleep(100);
  }
}

// ---------------------------------------------------------------------------------------------
// One exact item: slice `ch` (of S) of the deciding row pair of sequence b -- tail_item() of tail_fused.cuh with the
// inter-CTA exchange done through self-validating 8-byte words instead of atomics + fences + counters: every word the
// CTAs of a sequence exchange (slice normalisers, per-256-element residual partial sums, greedy keys) carries
// WORD_VALID in bit 63 and lives in memory zeroed at the head of the call, so a reader simply polls the words it needs;
// nobody executes a fence or waits for an atomic's return value.  The last slice of the sequence finalizes.
// ---------------------------------------------------------------------------------------------
// cold paths of the exact items, kept out of line so that they do not cost the hot loops registers
template <int DT>
__device__ __forceinline__ void mega_finalize(const DecideJob& job, const HybridWs& ws, int b, TailSh* sh) {
  finalize_sequence<DT>(job, ws, b, sh->sh64, sh->shf, sh->shi, &sh->s_res);
}
template <int DT>
__device__ __forceinline__ void mega_decide(const DecideJob& job, const HybridWs& ws, int b) {
  decide_sequence<DT>(job, ws, b);
}

// warp-collective: polls words w0 (and w0 + 1 if TWO) of all S slots of sequence b; sums (or maxima) of the payloads
template <bool TWO, bool MAXOP>
__device__ __forceinline__ bool slots_collect(const HybridWs& ws, int b, int S, int w0, u64& out0, u64& out1) {
  const int lane = threadIdx.x & 31;
  const u64* base = ws.xs + (size_t)b * ws.xs_stride * 4 + w0;
  u64 a0 = 0, a1 = 0;
  for (unsigned it = 1;; ++it) {
    bool ok = true;
    a0 = 0; a1 = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      if (s < S) {
        const u64 v0 = ld_relaxed_u64(base + (size_t)s * 4), v1 = TWO ? ld_relaxed_u64(base + (size_t)s * 4 + 1) : WORD_VALID;
        ok = ok && ((v0 & v1 & WORD_VALID) != 0ull);
        const u64 p0 = v0 & ~WORD_VALID, p1 = v1 & ~WORD_VALID;
        if (MAXOP) { a0 = p0 > a0 ? p0 : a0; } else { a0 += p0; a1 += p1; }
      }
    }
    if (__all_sync(0xffffffffu, ok)) break;
    if ((it & 255u) == 0u) {
      int ab = 0;
      if (lane == 0) {
        ab = ld_acquire_gpu(ws.abort);
        if (!ab && it > (1u << 22)) { atomicExch(ws.abort, 1); ab = 1; }
      }
      if (__shfl_sync(0xffffffffu, ab, 0)) return false;
    }
    __nanosleep(SPECDEC_POLL_NS);
  }
  if (MAXOP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, a0, o); a0 = t > a0 ? t : a0; }
  } else {
    a0 = warp_sum_u64(a0);
    if (TWO) a1 = warp_sum_u64(a1);
  }
  out0 = a0; out1 = a1;
  return true;
}

template <int DT, bool GREEDY, int NS_ITEM = MG_NS>
__device__ __forceinline__ bool mega_item(const DecideJob& job, const HybridWs& ws, const int b, const int ch, const int S,
                                          const int segs_per_cta, float4* ecache, TailSh& sh, const int seq_tasks, int4 rec,
                                          int mcq_bits, volatile int* credit) {
  const RowJob& rj = job.rj;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = job.gamma, V = rj.V, rps = rj.nT + rj.nD;
  const int NV = (V + 7) >> 3, nseg = (NV + 31) >> 5;
  const int s0 = min(nseg, ch * segs_per_cta), s1 = min(nseg, s0 + segs_per_cta);
  const float c = rj.c;
  u64* slot = ws.xs + ((size_t)b * ws.xs_stride + ch) * 4;

  // ---- rare: exact sums of the ambiguous positions, then slice 0 decides (atomics + flags, as tail_item) ----
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
      if (ch == 0) sh.sh_ok = spin_until(&ws.exact_done[b], S, ws.abort) ? 1 : 0;
    }
    cta_sync();
    if (!sh.sh_ok) return false;
    if (ch == 0 && tid < 32) {
      mega_decide<DT>(job, ws, b);
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicExch(&ws.decided[b], 1);
      }
    }
    if (tid == 0) sh.sh_ok = spin_until(&ws.decided[b], 1, ws.abort) ? 1 : 0;
    cta_sync();
    if (!sh.sh_ok) return false;
    rec = __ldcg((const int4*)(ws.samp + b * SAMP_N));
    mcq_bits = __ldcg(&ws.samp[b * SAMP_N + 4]);
  }
  const int mode = rec.y, prow_i = rec.z;
  const float mcp = __int_as_float(rec.w), mcq = __int_as_float(mcq_bits);
  if (mode != 2 && tid == 0) *credit = 1;  // (no sibling normalisers to wait for)
  if (mode == 0) return true;

  const long long r1 = (long long)b * rps + prow_i, r2 = (long long)b * rps + rj.nT + prow_i;
  const void* prowp = seq_row_ptr<DT>(rj, b, prow_i);
  const bool pal = (((size_t)prowp) & 15) == 0;
  u64* part = ws.part + (size_t)b * ws.nseg_pad;
  u64 tot = 0, Sp = 0, Sq = 0;
  float best = (mode == 2) ? 0.0f : -1.0f;
  int bidx = 0x7FFFFFFF;
  if (mode == 1) {
    // bonus / target row itself: the sampling weights are the canonical weights, one evaluation suffices
    RowOut rp;
    rp.m = 0.0f; rp.mc = mcp; rp.inv = 0.0f; rp.cut = -INFINITY; rp.jcut = V; rp.flags = 0; rp.Sfix = 0;
    partial_loop<DT, false, GREEDY, false>(prowp, prowp, rp, rp, pal, pal, V, c, s0, s1, part, tot, best, bidx, WORD_VALID);
  } else {
    const void* qrowp = seq_row_ptr<DT>(rj, b, rj.nT + prow_i);
    const bool qal = (((size_t)qrowp) & 15) == 0;
    // ---- phase A: canonical weights of my slice -> shared memory; partial normalisers -> my slot ----
    u64 sp = 0, sq = 0;
    pair_sums<DT, true, NS_ITEM>(prowp, qrowp, pal, qal, V, c, mcp, mcq, s0, s1, ecache, sp, sq);
    block_sum2_u64(sp, sq, sh.sh64);
    if (tid == 0) {
      dbg_stamp_max(ws, 16 + b * 8 + 7);
      st_relaxed_u64(slot, sp | WORD_VALID);
      st_relaxed_u64(slot + 1, sq | WORD_VALID);
    }
    if (w == 0) {
      const bool ok = slots_collect<true, false>(ws, b, S, 0, Sp, Sq);
      if (lane == 0) { sh.sh_S[0] = Sp; sh.sh_S[1] = Sq; sh.sh_ok = ok ? 1 : 0; dbg_stamp_max(ws, 16 + b * 8 + 4); }
    }
    cta_sync();
    if (!sh.sh_ok) return false;
    if (tid == 0) *credit = 1;  // past the last wait for sibling slices (a finalizer only waits for LOWER items)
    Sp = sh.sh_S[0]; Sq = sh.sh_S[1];
    const float invp = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(Sp), 0x1p-40f));
    const float invq = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(Sq), 0x1p-40f));
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
      if (lane == 0) st_relaxed_u64(&part[seg], s | WORD_VALID);
    }
  }
  if (GREEDY) {  // (value, smallest index) maximum of my slice -> word 2 of my slot
    u64 key = (bidx == 0x7FFFFFFF) ? 0ull : (((u64)__float_as_uint(best)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)bidx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, key, o); key = t > key ? t : key; }
    if (lane == 0) sh.sh64[w] = key;
    cta_sync();
    if (tid == 0) {
      u64 k2 = 0;
      for (int i = 0; i < TF_T / 32; ++i) k2 = sh.sh64[i] > k2 ? sh.sh64[i] : k2;
      st_relaxed_u64(slot + 2, k2 | WORD_VALID);
    }
  }
  if (ch != S - 1) return true;
  // ---- last slice of the sequence: wait for every partial sum, then finalize ----
  cta_sync();  // (my own partial sums are written)
  if (w == 0) {
    u64 total = 0;
    bool ok = true;
    for (unsigned it = 1;; ++it) {
      bool all = true;
      total = 0;
      for (int i0 = 0; i0 < nseg; i0 += 128) {
        u64 pv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int i = i0 + q * 32 + lane; pv[q] = (i < nseg) ? ld_relaxed_u64(&part[i]) : WORD_VALID; }
#pragma unroll
        for (int q = 0; q < 4; ++q) { all = all && ((pv[q] & WORD_VALID) != 0ull); total += pv[q] & ~WORD_VALID; }
      }
      if (__all_sync(0xffffffffu, all)) break;
      if ((it & 255u) == 0u) {
        int ab = 0;
        if (lane == 0) {
          ab = ld_acquire_gpu(ws.abort);
          if (!ab && it > (1u << 22)) { atomicExch(ws.abort, 1); ab = 1; }
        }
        if (__shfl_sync(0xffffffffu, ab, 0)) { ok = false; break; }
      }
      __nanosleep(SPECDEC_POLL_NS);
    }
    total = warp_sum_u64(total);
    u64 bk = 0, dummy = 0;
    if (GREEDY && ok) ok = slots_collect<false, true>(ws, b, S, 2, bk, dummy);
    if (lane == 0) {
      // finalize_sequence() reads these through L2 (__ldcg): stored by this CTA, read after the barrier below
      ws.tot[b] = total;
      ws.best[b] = bk;
      if (mode == 2) { ws.acc[r1] = Sp; ws.acc[r2] = Sq; }
      sh.sh_ok = ok ? 1 : 0;
      dbg_stamp_max(ws, 16 + b * 8 + 5);
    }
  }
  cta_sync();
  if (!sh.sh_ok) return false;
  mega_finalize<DT>(job, ws, b, &sh);
  if (tid == 0) dbg_stamp_max(ws, 16 + b * 8 + 6);
  return true;
}

// ---------------------------------------------------------------------------------------------
// tail_slots_kernel: the fused tail of the three-launch pipeline (tail_fused_kernel's role and launch geometry) with
// mega_item's exchange.  A tail CTA's life is a chain of dependent L2 round trips, not arithmetic (~2 us of issue
// in a ~13 us life at 4 CTAs / SM: profiles/README.md): ticket, plan record, row loads, {2 atomics, fence, counter
// atomic, poll, 2 loads} for the normalisers, {atomic, fence, counter atomic with return} for the hand-over to the
// finalizer.  Self-validating words cut the two exchanges to {2 stores | poll} and {stores}: no fence, no atomic
// whose return value is waited for, 19 of 20 CTAs leave right after their last store.
// ---------------------------------------------------------------------------------------------
template <int DT, bool GREEDY>
__global__ void __launch_bounds__(TF_T, 4) tail_slots_kernel(DecideJob job, HybridWs ws, int segs_per_cta, int TF_CH) {
  extern __shared__ __align__(16) float4 ecache[];
  __shared__ TailSh sh;
  __shared__ int sh_ticket;
  __shared__ volatile int credit;
  if (threadIdx.x == 0) sh_ticket = atomicAdd(ws.ticket, 1);  // (taken before the dependency wait, see tail_fused_kernel)
  grid_dependency_wait();
  __syncthreads();
  const int b = sh_ticket / TF_CH, ch = sh_ticket - b * TF_CH;
  // the sequence's whole plan record in one round trip: {n, mode, p-row position, mc_p | mc_q, exact tasks, -, -}
  const int4 rec = __ldcg((const int4*)(ws.samp + b * SAMP_N)), rec2 = __ldcg((const int4*)(ws.samp + b * SAMP_N) + 1);
  mega_item<DT, GREEDY, 4>(job, ws, b, ch, TF_CH, segs_per_cta, ecache, sh, rec2.y, rec, rec2.x, &credit);
}

// ---------------------------------------------------------------------------------------------
// X role, compute group (warps 0-7): exact items from the mailbox
// ---------------------------------------------------------------------------------------------
template <int DT, bool GREEDY>
__device__ __forceinline__ void mega_exact_items(const DecideJob& job, const HybridWs& ws, const MegaCfg& cfg, MegaSh& sh,
                                                 float4* ecache) {
  const int tid = threadIdx.x;
  int slot = 0;
  for (;;) {
    if (tid == 0) {
      sh.tail.sh_ok = 1;
      for (unsigned it = 1; sh.slot_state[slot] == 0; ++it) {
        if ((it & 4095u) == 0u && ld_acquire_gpu(ws.abort) != 0) { sh.tail.sh_ok = 0; break; }
        __nanosleep(32);
      }
    }
    cta_sync();
    __threadfence_block();
    if (!sh.tail.sh_ok) break;
    const int item = sh.slot_item[slot];
    if (item < 0) break;
    int rec[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) rec[k] = sh.slot_rec[slot][k];
    cta_sync();
    if (tid == 0) sh.slot_state[slot] = 0;  // hand the slot back: the service warp claims the next item meanwhile
    slot ^= 1;
    const int b = item / cfg.S, s = item - b * cfg.S;
    if (tid == 0 && ws.dbg) { atomicMin(&ws.dbg[16 + b * 8 + 2], global_timer_ns()); atomicMax(&ws.dbg[16 + b * 8 + 3], global_timer_ns()); }
    if (!mega_item<DT, GREEDY>(job, ws, b, s, cfg.S, cfg.spc, ecache, sh.tail, rec[5],
                               make_int4(rec[0], rec[1], rec[2], rec[3]), rec[4], &sh.credit))
      break;
  }
}

template <int DT, bool GREEDY>
__global__ void __launch_bounds__(TS_THREADS, 4) verify_mega_kernel(DecideJob job, HybridWs ws, MegaCfg cfg) {
  extern __shared__ __align__(128) unsigned char mega_dyn[];
  __shared__ __align__(8) unsigned long long full_bar[TS_STAGES], empty_bar[TS_STAGES];
  __shared__ float sh_m[2][8], sh_s[2][8];
  __shared__ MegaSh msh;
  __shared__ int sh_lane;
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < TS_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], TS_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    msh.slot_state[0] = 0; msh.slot_state[1] = 0; msh.credit = 1;
  }
  // Roles by arrival order on the SM: the first r_per_sm CTAs of every SM stream rows, the others are X CTAs.  The
  // grid is exactly (CTAs that fit one SM) x (SMs) and all of them are resident (cooperative launch), so every SM
  // hosts the same mix; row-slice lanes are numbered by a global ticket.
  if (tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int k = atomicAdd(&ws.sm_slots[smid & (MG_SM_SLOTS - 1)], 1);
    sh_lane = (k < cfg.r_per_sm) ? atomicAdd(ws.r_ticket, 1) : -1 - atomicAdd(ws.r_ticket + 1, 1);
  }
  __syncthreads();
  const int r_lane = sh_lane;              // >= 0: R role, lane id;  < 0: X role, rank -1 - r_lane
  if (tid == 0) dbg_stamp_min(ws, 0);
  if (tid == 0 && ws.dbg && blockIdx.x < 1024) {  // which SM runs this CTA, and when its row streaming ended
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    ws.dbg[16 + cfg.B * 8 + blockIdx.x] = (u64)smid | (r_lane >= 0 ? 0x8000ull : 0ull);
  }
  if (r_lane >= 0 && r_lane < cfg.n_r) {
    mega_stream_rows<DT>(job, ws, cfg, r_lane, mega_dyn, full_bar, empty_bar, sh_m, sh_s);
    __syncthreads();  // every stage of the ring has been consumed: the 48 KB become this CTA's weight cache
    if (tid == 0) { dbg_stamp_max(ws, 1); dbg_stamp_min(ws, 3); }
    if (tid == 0 && ws.dbg && blockIdx.x < 1024) ws.dbg[16 + cfg.B * 8 + blockIdx.x] |= (global_timer_ns() - ws.dbg[0]) << 16;
  }
  if ((tid >> 5) == TS_CONSUMERS / 32) mega_service_warp<DT>(job, ws, cfg, msh, r_lane < 0 ? -1 - r_lane : -1);
  else mega_exact_items<DT, GREEDY>(job, ws, cfg, msh, reinterpret_cast<float4*>(mega_dyn));
  if (tid == 0) dbg_stamp_max(ws, 2);
}
