// hybrid.cuh -- the production verify pipeline (included by verify.cu inside namespace specdec).
//
// The canonical (oracle-reproducible) arithmetic costs ~12 issue slots per logit because exp2 is a
// polynomial on the FMA pipe.  Only a few rows per sequence actually need it, so the step is split:
//
//   1. rowfast_tma_kernel / rowfast_kernel: every row, ONE pass over HBM: online max + sum of MUFU
//      ex2 (5 slots/logit).  The row max is exact; the sum is good to ~1e-6 relative.
//   2. plan_kernel (one warp per sequence): p~/q~ of the draft tokens from the fast sums and the accept
//      test with a 1e-3 relative safety margin => sure-accept / sure-reject / ambiguous.  Sequences
//      without ambiguous positions are decided on the spot.
//   3. tail_fused_kernel (tail_fused.cuh, default): exact normalisers of the deciding row pair, residual
//      partial sums from the weights cached in shared memory, token location -- one launch.
//   3'/4'. exact_rows_kernel + sample_partial_kernel: the same work as two launches without the cache
//      (fallback for huge vocabularies, test hook "no_fused_tail"; sample_partial also serves the masked
//      modes).  Their tails are chained with "last CTA done" counters (no host sync, no spin); the fused tail's
//      CTAs of one sequence DO wait for one another (bounded, ticket-ordered: tail_fused.cuh) and the launcher falls
//      back to this pair whenever fewer CTAs than one sequence needs can be co-resident.
//
// top-k / nucleus modes use the exact rowstats_kernel for every row (their kept sets need exact
// masses) followed by plan_kernel (no tasks).  Outputs are bit-identical to the exact-everywhere path.
#pragma once

constexpr int FT = 512;   // threads, rowfast_kernel
constexpr int PT = 256;   // threads, chunk kernels
constexpr int CH = 8;     // CTAs per row (pair) in exact_rows / sample_partial
constexpr int ST_ACCEPT = 0, ST_REJECT = 1, ST_AMBIG = 2, ST_EXACTROW = 3, ST_NEED = 4;
constexpr float MARGIN = 1e-3f;
constexpr int SAMP_N = 8;  // ints per sequence in HybridWs::samp
constexpr int MG_MAXU = 8;  // megakernel: slices per logit row
constexpr int MG_SM_SLOTS = 256;  // megakernel: >= number of SM ids
constexpr int MG_MIN_SPC = 12;  // megakernel: smallest exact item (256-element segments); sizes the slot array

struct HybridWs {
  u64* acc;               // [R]  canonical Sfix of task rows
  int* tasks;             // [B*gamma]
  int* ntasks;            // [1]
  unsigned char* status;  // [B*gamma]
  u64* part;              // [B][nseg_pad]
  u64* tot;               // [B]
  u64* best;              // [B]  greedy: (value bits << 32) | ~index
  int* samp;              // [B][SAMP_N]: n, mode, p-row position, bits of mc of the p row, of the q row, exact tasks
  int* rows_done;         // [B]  rows of the sequence whose statistics are written
  int* seq_tasks;         // [B]  number of exact tasks of the sequence
  int* exact_done;        // [B]
  int* part_done;         // [B]
  u64* acc2;              // [B][2]  fused tail: canonical Sfix of the deciding row pair
  int* fin_done;          // [B]     fused tail: CTAs whose partial sums are in acc2
  int* decided;           // [B]     fused tail: set once CTA 0 of the group has decided the sequence
  int* ticket;            // [2]     fused tail: logical CTA ids in dispatch order (one counter per half batch)
  int* abort;             // [1]     set when a bounded inter-CTA wait gave up (see spin_until)
  int* r_claim;           // [1]     row kernel: next unclaimed row (one counter per chunk); nullptr = static rows
  int* plan_done;         // [B]     megakernel: the sequence's plan record is published
  int* x_next;            // [1]     megakernel: next exact item (sequence-major, claimed in order)
  int* p_next;            // [1]     megakernel: next sequence to plan
  float2* rpart;          // [R][MG_MAXU]  megakernel: (max, MUFU sum relative to it) of every row slice
  int* sm_slots;          // [MG_SM_SLOTS]  megakernel: CTAs that arrived per SM (role assignment)
  int* r_ticket;          // [2]     megakernel: R lane / X rank tickets
  u64* xs;                // [B][xs_stride][4]  megakernel: per exact item {Sfix_p part, Sfix_q part, greedy key, -} | WORD_VALID
  int xs_stride;
  u64* dbg;               // nullable: [16 + B * 8] %globaltimer stamps of the megakernel (option "mega_dbg")
  int fused;              // plan: the first sure reject is handled by tail_fused_kernel, not as an exact task
  int nseg_pad;
};


__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Bounded wait for *p >= target.  `abort` (a word in the zeroed part of the workspace) turns a wait that can never be
// satisfied -- sibling CTAs that were not co-scheduled (MPS / sanitizer / a debugger serialising CTAs) -- into an
// error instead of a hang: after ~2^22 polls the waiter sets it, every other waiter sees it and leaves, the kernel
// terminates and specdec_verify_status() reports SPECDEC_ERR_TIMEOUT.
__device__ __forceinline__ bool spin_until(const int* p, int target, int* abort) {
  for (unsigned it = 1;; ++it) {
    if (ld_acquire_gpu(p) >= target) return true;
    if ((it & 255u) == 0u && abort) {
      if (ld_acquire_gpu(abort) != 0) return false;
      if (it > (1u << 22)) { atomicExch(abort, 1); return false; }
    }
    __nanosleep(40);
  }
}
// L2-coherent reads of records other CTAs of the SAME launch wrote (a persistent CTA may hold stale L1 lines)
__device__ __forceinline__ RowOut ldcg_rowout(const RowOut* p) {
  const int4 a = __ldcg(reinterpret_cast<const int4*>(p)), b = __ldcg(reinterpret_cast<const int4*>(p) + 1);
  RowOut o;
  o.m = __int_as_float(a.x); o.mc = __int_as_float(a.y); o.inv = __int_as_float(a.z); o.cut = __int_as_float(a.w);
  o.jcut = b.x; o.flags = b.y;
  o.Sfix = ((u64)(unsigned)b.w << 32) | (u64)(unsigned)b.z;
  return o;
}

__device__ __forceinline__ u64 global_timer_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// debug timeline (HybridWs::dbg): global slots 0..15, then 8 per sequence
__device__ __forceinline__ void dbg_stamp_min(const HybridWs& ws, int slot) { if (ws.dbg) atomicMin(&ws.dbg[slot], global_timer_ns()); }
__device__ __forceinline__ void dbg_stamp_max(const HybridWs& ws, int slot) { if (ws.dbg) atomicMax(&ws.dbg[slot], global_timer_ns()); }

__device__ __forceinline__ float job_u_accept(const DecideJob& job, int b, int i) {
  return job.u_accept ? job.u_accept[(long long)b * job.gamma + i]
                      : philox_uniform(job.seed, job_offset(job), (unsigned)(job.seq0 + b), (unsigned)i);
}
__device__ __forceinline__ int accept_rule(float p, float q, float u, int flags) {
  if (flags & SPECDEC_ACCEPT_BATCHED) {  // engine/infer_engine.py:303-305 (python floats = double)
    const double ap = (q <= 0.0f) ? 1.0 : fmin(1.0, (double)p / (double)q);
    return ((double)u < ap) ? 1 : 0;
  }
  const float frac = __fdiv_rn(p, q);  // sampling/speculative_decoding.py:143 (NaN => accept)
  return !(u > frac) ? 1 : 0;
}

// exact statistics of row r: from rowstats_kernel (masked modes) or from the exact task sums
__device__ __forceinline__ RowOut resolved_row(const RowJob& rj, const HybridWs& ws, long long r) {
  RowOut o = ldcg_rowout(&rj.out[r]);
  if (!(o.flags & 1)) {
    const u64 S = __ldcg(&ws.acc[r]);
    o.Sfix = S;
    o.inv = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(S), 0x1p-40f));
  }
  return o;
}

// ---------------------------------------------------------------------------------------------
// decide: executed by ONE WARP once every exact sum the sequence needs is in ws.acc
// ---------------------------------------------------------------------------------------------
template <int DT>
__device__ void decide_sequence(const DecideJob& job, const HybridWs& ws, int b) {
  const RowJob& rj = job.rj;
  const int g = job.gamma, lane = threadIdx.x & 31;
  const long long* toks = job.draft_tokens + (long long)b * g;
  int n = g;
  for (int i0 = 0; i0 < g; i0 += 32) {
    const int i = i0 + lane;
    int a = 1;
    if (i < g) {
      const int st = __ldcg(&ws.status[(long long)b * g + i]);  // (possibly written by another CTA of this launch)
      if ((st & ST_NEED) || (st & 3) == ST_EXACTROW) {
        const int rps = rj.nT + rj.nD;
        const int tok = (int)min(max(toks[i], 0ll), (long long)rj.V - 1);
        const long long r1 = (long long)b * rps + i, r2 = (long long)b * rps + rj.nT + i;
        const RowOut rp = resolved_row(rj, ws, r1), rq = resolved_row(rj, ws, r2);
        const float p = row_prob<DT>(rp, row_ptr<DT>(rj, r1), tok, rj.c);
        const float q = row_prob<DT>(rq, row_ptr<DT>(rj, r2), tok, rj.c);
        a = accept_rule(p, q, job_u_accept(job, b, i), job.flags);
        job.p_tok[(long long)b * g + i] = p;
        job.q_tok[(long long)b * g + i] = q;
      } else {
        a = ((st & 3) == ST_ACCEPT);
      }
      job.mask[(long long)b * g + i] = (unsigned char)a;
    }
    const unsigned rej = __ballot_sync(0xffffffffu, i < g && !a);
    if (n == g && rej) n = i0 + __ffs(rej) - 1;
  }
  if (lane == 0) {
    int mode = 0, prow = 0;  // 0 none, 1 target row, 2 residual
    if (n == g) {
      if (!(job.flags & SPECDEC_NO_BONUS)) { mode = 1; prow = g; }
    } else if (job.flags & SPECDEC_SKIP_ADJUST) { mode = 1; prow = n; }
    else { mode = 2; prow = n; }
    const int fs = first_stop_index(toks, n, job.stop, job.n_stop, job.flags);
    job.n_acc[b] = n;
    job.first_stop[b] = fs;
    ws.samp[b * SAMP_N + 0] = n; ws.samp[b * SAMP_N + 1] = mode; ws.samp[b * SAMP_N + 2] = prow;
    if (mode) {  // one record read gives tail_fused_kernel everything it needs to start streaming
      const int rps = rj.nT + rj.nD;
      ws.samp[b * SAMP_N + 3] = __float_as_int(ldcg_rowout(&rj.out[(long long)b * rps + prow]).mc);
      ws.samp[b * SAMP_N + 4] = (mode == 2) ? __float_as_int(ldcg_rowout(&rj.out[(long long)b * rps + rj.nT + prow]).mc) : 0;
    }
    if (mode == 0) {  // nothing to sample (all accepted, no bonus token)
      job.next_tok[b] = -1;
      if (job.next_prob) job.next_prob[b] = 0.0f;
      if (job.packed) {
        int* pk = job.packed + (long long)b * (g + 2);
        pk[0] = n;
        for (int i = 0; i < g; ++i) pk[1 + i] = (int)toks[i];
        pk[1 + g] = -1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plan: executed by ONE WARP once all RowOut of the sequence are written (lane = draft position)
// ---------------------------------------------------------------------------------------------
template <int DT>
__device__ void plan_sequence(const DecideJob& job, const HybridWs& ws, int b) {
  const int lane = threadIdx.x & 31;
  const RowJob& rj = job.rj;
  const int g = job.gamma, rps = rj.nT + rj.nD;
  const RowOut* ro = rj.out + (long long)b * rps;
  const long long* toks = job.draft_tokens + (long long)b * g;
  bool have_reject = false;
  int ntask = 0;
  for (int i0 = 0; i0 < g; i0 += 32) {
    const int i = i0 + lane;
    int st = ST_ACCEPT;
    bool exact_row = false;
    if (i < g) {
      const RowOut rp = ro[i];
      const RowOut rq = ro[rj.nT + i];
      if (rp.flags & 1) {
        st = ST_EXACTROW; exact_row = true;
      } else {
        const int tok = (int)min(max(toks[i], 0ll), (long long)rj.V - 1);
        const float zp = load1<DT>(row_ptr<DT>(rj, (long long)b * rps + i), tok);
        const float zq = load1<DT>(row_ptr<DT>(rj, (long long)b * rps + rj.nT + i), tok);
        const float p = exp2f(__fmaf_rn(zp, rj.c, -rp.mc)) * rp.inv;
        const float q = exp2f(__fmaf_rn(zq, rj.c, -rq.mc)) * rq.inv;
        job.p_tok[(long long)b * g + i] = p;
        job.q_tok[(long long)b * g + i] = q;
        const float u = job_u_accept(job, b, i);
        st = ST_AMBIG;
        if (p > 1e-30f && q > 1e-30f && p < 1e30f && q < 1e30f) {
          const float r = p / q, lo = r * (1.0f - MARGIN), hi = r * (1.0f + MARGIN);
          if (job.flags & SPECDEC_ACCEPT_BATCHED) {
            if (u < fminf(1.0f, lo)) st = ST_ACCEPT;
            else if (hi < 1.0f && u > hi) st = ST_REJECT;
          } else {
            if (u < lo) st = ST_ACCEPT;
            else if (u > hi) st = ST_REJECT;
          }
        }
      }
    }
    const unsigned rej = __ballot_sync(0xffffffffu, i < g && !exact_row && st == ST_REJECT);
    bool need = (i < g) && !exact_row && (st == ST_AMBIG);
    if (!have_reject && rej) {
      if (!ws.fused && lane == __ffs(rej) - 1) need = true;
      have_reject = true;
    }
    if (i < g) {
      if (need) {
        st |= ST_NEED;
        if (!ws.fused) {
          const int t = atomicAdd(ws.ntasks, 1);
          ws.tasks[t] = b * g + i;
        }
      }
      ws.status[(long long)b * g + i] = (unsigned char)st;
    }
    ntask += __popc(__ballot_sync(0xffffffffu, need));
  }
  if (lane == 0) { ws.seq_tasks[b] = ntask; ws.samp[b * SAMP_N + 5] = ntask; }
  __syncwarp();
  if (ntask == 0) {  // nothing to make exact: decide right away
    __threadfence();
    decide_sequence<DT>(job, ws, b);
  }
}

// one warp per sequence (masked modes / gamma == 0, where no fast row kernel runs)
// Programmatic dependent launch (PDL): plan_kernel and tail_fused_kernel are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so their CTAs are scheduled while the previous kernel of the
// stream drains; griddepcontrol.wait blocks until that kernel has completed and its writes are visible (a no-op
// for a normal launch).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <int DT>
__global__ void __launch_bounds__(256) plan_kernel(DecideJob job, HybridWs ws) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the tail's CTAs may start taking their tickets
  grid_dependency_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= (int)(job.rj.R / (job.rj.nT + job.rj.nD))) return;
  plan_sequence<DT>(job, ws, b);
}

// ---------------------------------------------------------------------------------------------
// row kernel, vectorised-LDG version (fp32 rows, unaligned rows)
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(FT, 2) rowfast_kernel(DecideJob dj, HybridWs ws) {
  __shared__ float shf[33];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // PDL: plan_kernel may be scheduled early (it waits)
  const RowJob& job = dj.rj;
  const long long r = blockIdx.x;
  const void* row = row_ptr<DT>(job, r);
  const bool aligned = (((size_t)row) & 15) == 0;
  const int V = job.V, NV = (V + 7) >> 3;
  const float c = job.c;
  float m = -INFINITY, s = 0.0f;
  sweep_range<DT, FT>(row, V, aligned, 0, NV, [&](const float(&x)[8], int) {
    float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
    if (vm > m) {  // rare after the first few vectors
      s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, vm), c)));
      m = vm;
    }
    const float mc = (m > -INFINITY) ? __fmul_rn(m, c) : 0.0f;  // (only -inf so far: -inf*c + inf would be NaN)
#pragma unroll
    for (int k = 0; k < 8; ++k) s = __fadd_rn(s, ex2_approx(__fmaf_rn(x[k], c, -mc)));
  });
  const float M = block_max_f(m, shf);
  s = (m > -INFINITY) ? __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, M), c))) : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) shf[w] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (lane < FT / 32) ? shf[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      RowOut o;
      o.m = M; o.mc = __fmul_rn(M, c); o.inv = __fdiv_rn(1.0f, t);
      o.cut = -INFINITY; o.jcut = V; o.flags = 0; o.Sfix = 0;
      job.out[r] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// n-gram-assisted verify, greedy processor (ngram_assisted/ngram_assisted.py:114-141 with GreedyProcessor):
// every decision is "draft == argmax(p_i)" and the next token is argmax(p_n), so no normaliser is needed
// for the decisions at all.  rowfast_argmax_kernel = the fast row kernel + (first index of the row
// maximum, second largest distinct value).  The canonical arg-max is over the weights e_j = cexp2(t_j);
// it equals the arg-max over the logits whenever the runner-up is >= 4e-6 below the maximum in exponent
// units (20x the polynomial's error) -- always for bf16/fp16 logits; rows that are closer are redone
// exactly by ngram_greedy_decide_kernel.
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(FT, 2) rowfast_argmax_kernel(RowJob job) {
  __shared__ float shf[33];
  __shared__ unsigned shu[33];
  const long long r = blockIdx.x;
  const void* row = row_ptr<DT>(job, r);
  const bool aligned = (((size_t)row) & 15) == 0;
  const int V = job.V, NV = (V + 7) >> 3;
  const float c = job.c;
  float m = -INFINITY, s = 0.0f, s2 = -INFINITY;
  int idx = 0x7FFFFFFF;
  sweep_range<DT, FT>(row, V, aligned, 0, NV, [&](const float(&x)[8], int j0) {
    const float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
    if (vm > s2) {  // rare after the first few vectors
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xv = x[k];
        if (xv > m) {
          s = __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, xv), c)));
          s2 = m; m = xv; idx = j0 + k;
        } else if (xv < m && xv > s2) {
          s2 = xv;
        }
      }
    }
    const float mc = (m > -INFINITY) ? __fmul_rn(m, c) : 0.0f;  // (only -inf so far: -inf*c + inf would be NaN)
#pragma unroll
    for (int k = 0; k < 8; ++k) s = __fadd_rn(s, ex2_approx(__fmaf_rn(x[k], c, -mc)));
  });
  const float M = block_max_f(m, shf);
  const unsigned first = block_min_u32((m == M) ? (unsigned)idx : 0xFFFFFFFFu, shu);
  const float second = block_max_f((m < M) ? m : s2, shf);
  s = (m > -INFINITY) ? __fmul_rn(s, ex2_approx(__fmul_rn(__fsub_rn(m, M), c))) : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) shf[w] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (lane < FT / 32) ? shf[lane] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      const bool amb = !(M > -INFINITY) || !(M < INFINITY) || first >= (unsigned)V ||
                       (second > -INFINITY && !(__fmul_rn(__fsub_rn(M, second), c) >= 4e-6f));
      RowOut o;
      o.m = M; o.mc = __fmul_rn(M, c); o.inv = __fdiv_rn(1.0f, t);
      o.cut = -INFINITY; o.jcut = V; o.flags = 0;
      o.Sfix = (u64)first | (amb ? (1ull << 32) : 0ull);
      job.out[r] = o;
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(PT, 4) ngram_greedy_decide_kernel(DecideJob job, HybridWs ws) {
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ int shi[33];
  __shared__ long long s_res;
  __shared__ long long s_tok[66];
  const RowJob& rj = job.rj;
  const int b = blockIdx.x, g = job.gamma, V = rj.V, nT = rj.nT;
  const RowOut* ro = rj.out + (long long)b * nT;
  const Scratch scr{ws.part + (size_t)b * ws.nseg_pad, sh64, shf, shi, &s_res};
  for (int i = threadIdx.x; i < nT; i += blockDim.x) {
    const u64 v = ro[i].Sfix;
    s_tok[i] = ((v >> 32) & 1ull) ? -1ll : (long long)(v & 0xFFFFFFFFull);
  }
  __syncthreads();
  for (int i = 0; i < nT; ++i) {  // block-uniform: exact arg-max for the (rare) near-tie rows
    if (s_tok[i] < 0) {
      const long long x = sample_p_row<DT>(row_ptr<DT>(rj, (long long)b * nT + i), ro[i], V, rj.c, true, 0.0f, scr);
      if (threadIdx.x == 0) s_tok[i] = x;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const long long* toks = job.draft_tokens + (long long)b * g;
    int n = g;
    for (int i = 0; i < g; ++i) {
      const int tok = (int)min(max(toks[i], 0ll), (long long)V - 1);
      const int a = (s_tok[i] == toks[i]) ? 1 : 0;
      job.mask[(long long)b * g + i] = (unsigned char)a;
      const float z = load1<DT>(row_ptr<DT>(rj, (long long)b * nT + i), tok);
      job.p_tok[(long long)b * g + i] = exp2f(fmaxf(__fmaf_rn(z, rj.c, -ro[i].mc), -125.0f)) * ro[i].inv;  // (canonical clamp: -inf logits)
      job.q_tok[(long long)b * g + i] = 0.0f;
      if (!a && n == g) n = i;
    }
    const int fs = first_stop_index(toks, n, job.stop, job.n_stop, job.flags);
    const long long x = (n < nT) ? s_tok[n] : -1ll;  // argmax(p_n), or the bonus row; -1 without a bonus row
    job.n_acc[b] = n;
    job.first_stop[b] = fs;
    job.next_tok[b] = x;
    if (job.next_prob) {
      float np = 0.0f;
      if (x >= 0) np = exp2f(fmaxf(__fmaf_rn(load1<DT>(row_ptr<DT>(rj, (long long)b * nT + n), (int)x), rj.c, -ro[n].mc), -125.0f)) * ro[n].inv;
      job.next_prob[b] = np;
    }
    if (job.packed) {
      int* pk = job.packed + (long long)b * (g + 2);
      pk[0] = n;
      for (int i = 0; i < g + 1; ++i) pk[1 + i] = -1;
      for (int i = 0; i < n; ++i) pk[1 + i] = (int)toks[i];
      pk[1 + n] = (int)x;
    }
  }
}

#include "rowfast_tma.cuh"

// ---------------------------------------------------------------------------------------------
// canonical sums of the task rows: grid (B, CH), tasks looped over; tail = decide
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(PT, 4) exact_rows_kernel(DecideJob job, HybridWs ws) {
  __shared__ u64 sh64[33];
  __shared__ int sh_last;
  grid_dependency_wait();
  const RowJob& rj = job.rj;
  const int ntasks = *ws.ntasks;
  for (int tix = blockIdx.x; tix < ntasks; tix += gridDim.x) {  // usually one task per sequence
  const int task = ws.tasks[tix];
  const int g = job.gamma, rps = rj.nT + rj.nD, V = rj.V;
  const int b = task / g, i = task - b * g;
  const int NV = (V + 7) >> 3, per = (NV + CH - 1) / CH;
  const int v0 = blockIdx.y * per, v1 = min(NV, v0 + per);
  const float c = rj.c;
  // both rows of the position in one loop: 8 independent 16-byte loads in flight per thread
  const long long r1 = (long long)b * rps + i, r2 = (long long)b * rps + rj.nT + i;
  const void* prow = row_ptr<DT>(rj, r1);
  const void* qrow = row_ptr<DT>(rj, r2);
  const bool pal = (((size_t)prow) & 15) == 0, qal = (((size_t)qrow) & 15) == 0;
  const float mcp = rj.out[r1].mc, mcq = rj.out[r2].mc;
  const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mcp, -mcq);
  u64 sp = 0, sq = 0;
  constexpr int NQ = (DT == DT_F32) ? 2 : 4;  // vectors of each row in flight per thread (raw, still packed)
  for (int v = v0 + threadIdx.x; v < v1; v += NQ * PT) {
    Raw8<DT> rp[NQ], rq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int vv = min(v + q * PT, v1 - 1);
      rp[q] = load_raw8<DT>(prow, vv, V, pal);
      rq[q] = load_raw8<DT>(qrow, vv, V, qal);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (v + q * PT < v1) {
        float xp[8], xq[8];
        unpack8<DT>(rp[q], xp);
        unpack8<DT>(rq[q], xq);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // (target, drafter) logit of the same token as one fp32x2 pair
          const float2 e = cweight2(make_float2(xp[k], xq[k]), c2, nmc2);
          sp += fix40(e.x);
          sq += fix40(e.y);
        }
      }
    }
  }
  sp = block_sum_u64(sp, sh64);
  sq = block_sum_u64(sq, sh64);
  if (threadIdx.x == 0) {
    if (sp) atomicAdd(&ws.acc[r1], sp);
    if (sq) atomicAdd(&ws.acc[r2], sq);
  }
  if (threadIdx.x == 0) {
    __threadfence();
    sh_last = (atomicAdd(&ws.exact_done[b], 1) == __ldcg(&ws.seq_tasks[b]) * CH - 1) ? 1 : 0;
  }
  __syncthreads();
  if (sh_last && threadIdx.x < 32) {
    __threadfence();
    decide_sequence<DT>(job, ws, b);
  }
  __syncthreads();
  }
}

struct SampleCtx {
  const void* prow;
  const void* qrow;
  RowOut rp, rq;
  bool pal, qal;
  int V;
  float c;
  int resid;
};
// integer weights of vector v (8 elements); r_out (nullable) receives the fp32 residuals / weights
template <int DT>
__device__ __forceinline__ void sample_weights(const SampleCtx& sc, int v, u64 (&w)[8], float* vals) {
  float xp[8];
  load8<DT>(sc.prow, v, sc.V, sc.pal, xp);
  if (sc.resid) {
    float xq[8];
    load8<DT>(sc.qrow, v, sc.V, sc.qal, xq);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = v * 8 + k;
      float r = 0.0f;
      if (j < sc.V) {
        const float P = __fmul_rn(kept(sc.rp, xp[k], j) ? cweight(xp[k], sc.c, sc.rp.mc) : 0.0f, sc.rp.inv);
        const float Q = __fmul_rn(kept(sc.rq, xq[k], j) ? cweight(xq[k], sc.c, sc.rq.mc) : 0.0f, sc.rq.inv);
        r = __fsub_rn(P, Q);
        r = r > 0.0f ? r : 0.0f;
      }
      w[k] = fix60(r);
      if (vals) vals[k] = r;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = v * 8 + k;
      const float e = (j < sc.V && kept(sc.rp, xp[k], j)) ? cweight(xp[k], sc.c, sc.rp.mc) : 0.0f;
      w[k] = fix40(e);
      if (vals) vals[k] = e;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// finalize: block-wide, by the last of the CH CTAs of a sequence
// ---------------------------------------------------------------------------------------------
template <int DT>
__device__ void finalize_sequence(const DecideJob& job, const HybridWs& ws, int b, u64* sh64, float* shf, int* shi,
                                  long long* s_res) {
  const RowJob& rj = job.rj;
  const int g = job.gamma, V = rj.V, rps = rj.nT + rj.nD;
  const int n = __ldcg(&ws.samp[b * SAMP_N + 0]), mode = __ldcg(&ws.samp[b * SAMP_N + 1]), prow = __ldcg(&ws.samp[b * SAMP_N + 2]);
  const bool greedy = job.greedy != 0;
  const float us = job.u_sample ? job.u_sample[b]
                                : philox_uniform(job.seed, job_offset(job), (unsigned)(job.seq0 + b), (unsigned)job.lane_sample);
  u64* part = ws.part + (size_t)b * ws.nseg_pad;
  const long long r1 = (long long)b * rps + prow;
  const void* prow_ptr = row_ptr<DT>(rj, r1);
  RowOut rp = resolved_row(rj, ws, r1);
  const u64 total = __ldcg(&ws.tot[b]);
  const u64 rmin = (job.flags & SPECDEC_RESID_FALLBACK) ? 1152921ull : 0ull;
  long long x;
  bool from_p;
  if (mode == 2 && total <= rmin) {
    // residual mass (numerically) zero: sample the target row itself (engine/infer_engine.py:319-321)
    const Scratch scr{part, sh64, shf, shi, s_res};
    x = sample_p_row<DT>(prow_ptr, rp, V, rj.c, greedy, us, scr);
    from_p = true;
  } else if (greedy) {
    const u64 key = __ldcg(&ws.best[b]);
    x = key ? (long long)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull)) : 0ll;
    from_p = (mode == 1);
  } else {
    SampleCtx sc;
    sc.prow = prow_ptr; sc.rp = rp; sc.pal = (((size_t)prow_ptr) & 15) == 0;
    sc.V = V; sc.c = rj.c; sc.resid = (mode == 2);
    sc.qrow = prow_ptr; sc.rq = rp; sc.qal = sc.pal;
    if (mode == 2) {
      const long long r2 = (long long)b * rps + rj.nT + prow;
      sc.qrow = row_ptr<DT>(rj, r2);
      sc.rq = resolved_row(rj, ws, r2);
      sc.qal = (((size_t)sc.qrow) & 15) == 0;
    }
    const int NV = (V + 7) >> 3, nseg = (NV + 31) >> 5;
    auto wf = [&](int v, u64(&w)[8]) { sample_weights<DT>(sc, v, w, nullptr); };
    x = locate_token(NV, nseg, part, total, scale_u24(total, u24_of(us)), wf, s_res);
    from_p = (mode == 1);
  }
  if (threadIdx.x == 0) {
    job.next_tok[b] = x;
    if (job.next_prob) {
      float np = 0.0f;
      if (from_p && x >= 0) {
        if (mode == 1 && !(rp.flags & 1))  // target row sampled without a prior exact sum: total IS its Sfix
          rp.inv = __fdiv_rn(1.0f, __fmul_rn(__ull2float_rn(total), 0x1p-40f));
        np = row_prob<DT>(rp, prow_ptr, (int)x, rj.c);
      }
      job.next_prob[b] = np;
    }
    if (job.packed) {
      const long long* toks = job.draft_tokens + (long long)b * g;
      int* pk = job.packed + (long long)b * (g + 2);
      pk[0] = n;
      for (int i = 0; i < g + 1; ++i) pk[1 + i] = -1;
      for (int i = 0; i < n; ++i) pk[1 + i] = (int)toks[i];
      pk[1 + n] = (int)x;
    }
  }
}

// per-256-element integer partial sums of the sampling weights of segments [s0, s1) (one warp per
// segment, two segments in flight).  RESID: weights fix60(max(0, P - Q)); else fix40(e) of the target row.
// Elements past V read as -inf => weight exactly 0, so only the greedy arg-max needs a bounds check.
template <int DT, bool MASKED, bool GREEDY, bool RESID>
__device__ __forceinline__ void partial_loop(const void* prowp, const void* qrowp, const RowOut& rp, const RowOut& rq,
                                             bool pal, bool qal, int V, float c, int s0, int s1, u64* part, u64& tot,
                                             float& best, int& bidx, const u64 vflag = 0ull) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int NV = (V + 7) >> 3;
  const float mcp = rp.mc, mcq = rq.mc, invp = rp.inv, invq = rq.inv;
  const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mcp, -mcq), inv2 = make_float2(invp, invq);
  auto seg_sum = [&](const float(&xp)[8], const float(&xq)[8], int v) -> u64 {
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = v * 8 + k;
      float val;
      if (RESID) {
        const bool kp = !MASKED || kept(rp, xp[k], j), kq = !MASKED || kept(rq, xq[k], j);
        float2 e = cweight2(make_float2(xp[k], xq[k]), c2, nmc2);
        if (MASKED) { e.x = kp ? e.x : 0.0f; e.y = kq ? e.y : 0.0f; }
        const float2 PQ = __fmul2_rn(e, inv2);
        val = fmaxf(__fsub_rn(PQ.x, PQ.y), 0.0f);
        s += fix60(val);
      } else {
        const bool kp = !MASKED || kept(rp, xp[k], j);
        val = kp ? cweight(xp[k], c, mcp) : 0.0f;
        s += fix40(val);
      }
      if (GREEDY && j < V && val > best) { best = val; bidx = j; }
    }
    return s;
  };
  constexpr int WPB = PT / 32;
  constexpr int NS = (DT == DT_F32) ? 2 : 4;  // segments in flight per warp (raw, still packed loads)
  for (int seg = s0 + w; seg < s1; seg += NS * WPB) {
    Raw8<DT> rpv[NS], rqv[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int vv = min((seg + q * WPB) * 32 + lane, NV - 1);
      rpv[q] = load_raw8<DT>(prowp, vv, V, pal);
      if (RESID) rqv[q] = load_raw8<DT>(qrowp, vv, V, qal);
    }
    u64 sums[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int sg = seg + q * WPB, vv = sg * 32 + lane;
      sums[q] = 0;
      if (sg < s1 && vv < NV) {
        float xp[8], xq[8];
        unpack8<DT>(rpv[q], xp);
        if (RESID) unpack8<DT>(rqv[q], xq);
        sums[q] = seg_sum(xp, xq, vv);
      }
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const int sg = seg + q * WPB;
      if (sg < s1) {  // warp-uniform
        const u64 sa = warp_sum_u64(sums[q]);
        if (lane == 0) { part[sg] = sa | vflag; tot += sa; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// grid (B, CH): integer partial sums of the sampling weights; tail = finalize
// ---------------------------------------------------------------------------------------------
template <int DT, bool MASKED, bool GREEDY>
__global__ void __launch_bounds__(PT, 4) sample_partial_kernel(DecideJob job, HybridWs ws) {
  __shared__ u64 sh64[33];
  __shared__ float shf[33];
  __shared__ int shi[33];
  __shared__ long long s_res;
  __shared__ int sh_last;
  grid_dependency_wait();
  const RowJob& rj = job.rj;
  const int b = blockIdx.x, ch = blockIdx.y, V = rj.V, rps = rj.nT + rj.nD;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int mode = ws.samp[b * SAMP_N + 1], prow = ws.samp[b * SAMP_N + 2];
  if (mode == 0) return;
  if (MASKED && ws.part_done[b] >= CH) return;  // drawn from the kept-token lists by sample_lists_kernel
  const long long r1 = (long long)b * rps + prow;
  const void* prowp = row_ptr<DT>(rj, r1);
  const RowOut rp = resolved_row(rj, ws, r1);
  const bool pal = (((size_t)prowp) & 15) == 0;
  const void* qrowp = prowp;
  RowOut rq = rp;
  bool qal = pal;
  if (mode == 2) {
    const long long r2 = (long long)b * rps + rj.nT + prow;
    qrowp = row_ptr<DT>(rj, r2);
    rq = resolved_row(rj, ws, r2);
    qal = (((size_t)qrowp) & 15) == 0;
  }
  const int NV = (V + 7) >> 3, nseg = (NV + 31) >> 5;
  const int per = (nseg + CH - 1) / CH;
  const int s0 = ch * per, s1 = min(nseg, s0 + per);
  u64 tot = 0;
  float best = (mode == 2) ? 0.0f : -1.0f;
  int bidx = 0x7FFFFFFF;
  u64* part = ws.part + (size_t)b * ws.nseg_pad;
  if (mode == 2)
    partial_loop<DT, MASKED, GREEDY, true>(prowp, qrowp, rp, rq, pal, qal, V, rj.c, s0, s1, part, tot, best, bidx);
  else
    partial_loop<DT, MASKED, GREEDY, false>(prowp, qrowp, rp, rq, pal, qal, V, rj.c, s0, s1, part, tot, best, bidx);
  tot = block_sum_u64(tot, sh64);
  if (GREEDY) {
    // (value, smallest index) max: non-negative floats order like their bit patterns
    u64 key = (bidx == 0x7FFFFFFF) ? 0ull : (((u64)__float_as_uint(best)) << 32) | (u64)(0xFFFFFFFFu - (unsigned)bidx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, key, o); key = t > key ? t : key; }
    if (lane == 0 && key) atomicMax(&ws.best[b], key);
  }
  __syncthreads();  // every warp's part[] stores and atomicMax are issued
  if (threadIdx.x == 0) {
    if (tot) atomicAdd(&ws.tot[b], tot);
    __threadfence();
    sh_last = (atomicAdd(&ws.part_done[b], 1) == CH - 1) ? 1 : 0;
  }
  __syncthreads();
  if (sh_last) {
    __threadfence();
    finalize_sequence<DT>(job, ws, b, sh64, shf, shi, &s_res);
  }
}

// ---------------------------------------------------------------------------------------------
// sample_lists_kernel (masked modes): one WARP per sequence.  When the rows the next token is drawn from carry a
// kept-token list (RowJob::klist: top-k, top-k + top-p, small nuclei -- at most KL_MAX tokens), the draw needs no
// sweep over the vocabulary: the residual max(0, P - Q) lives on the target row's kept tokens.  The same integers
// as sample_partial_kernel + finalize_sequence are summed (fix60 residuals / fix40 weights in index order), so the
// token is bit-identical.  Sequences without lists are left to sample_partial_kernel (ws.part_done[b] stays 0;
// handled sequences set it to CH so that kernel's CTAs return at once).
// ---------------------------------------------------------------------------------------------
constexpr int SL_WARPS = 8;
struct ListSh {  // per-warp scratch of sample_lists_sequence
  int j[KL_MAX];
  float p[KL_MAX];
  u64 w[KL_MAX];
  int qj[KL_MAX];
  float qz[KL_MAX];
};
// one WARP: the draw of sequence b from the kept-token lists (see sample_lists_kernel)
template <int DT>
__device__ __forceinline__ void sample_lists_sequence(const DecideJob& job, const HybridWs& ws, const int b, ListSh& lsh) {
  const int lane = threadIdx.x & 31;
  const RowJob& rj = job.rj;
  const int g = job.gamma, V = rj.V, rps = rj.nT + rj.nD;
  const int n = ws.samp[b * SAMP_N + 0], mode = ws.samp[b * SAMP_N + 1], prow = ws.samp[b * SAMP_N + 2];
  if (mode == 0) { if (lane == 0) ws.part_done[b] = CH; return; }  // decide_sequence wrote the outputs
  const long long r1 = (long long)b * rps + prow, r2 = (long long)b * rps + rj.nT + prow;
  const RowOut rp = rj.out[r1];
  RowOut rq = rp;
  if (mode == 2) rq = rj.out[r2];
  const int cp = (rp.flags >> 8) & 0xFF, cq = (mode == 2) ? ((rq.flags >> 8) & 0xFF) : 1;
  if (cp == 0 || cq == 0) return;  // no list: sample_partial_kernel does this sequence
  const float c = rj.c;
  const bool greedy = job.greedy != 0;
  const float us = job.u_sample ? job.u_sample[b]
                                : philox_uniform(job.seed, job_offset(job), (unsigned)(job.seq0 + b), (unsigned)job.lane_sample);
  int* sj = lsh.j; float* sp = lsh.p; u64* sw = lsh.w; int* sqj = lsh.qj; float* sqz = lsh.qz;
  // ---- load the lists; the target row's list sorted by token index (rank sort: KL_MAX^2 / 32 comparisons per lane)
  const int2* lp = rj.klist + r1 * KL_MAX;
  int2 e[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) e[h] = (lane + 32 * h < cp) ? lp[lane + 32 * h] : make_int2(0, 0x7FFFFFFF);
  if (mode == 2) {
    const int2* lq = rj.klist + r2 * KL_MAX;
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (lane + 32 * h < cq) { const int2 q = lq[lane + 32 * h]; sqj[lane + 32 * h] = q.y; sqz[lane + 32 * h] = __int_as_float(q.x); }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) sj[lane + 32 * h] = e[h].y;
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (lane + 32 * h < cp) {
      int rank = 0;
      for (int i = 0; i < cp; ++i) rank += (sj[i] < e[h].y) ? 1 : 0;
      // probabilities exactly as row_prob(): kept tokens only, canonical weight times the exact 1/S
      const float zp = __int_as_float(e[h].x);
      const float P = __fmul_rn(cweight(zp, c, rp.mc), rp.inv);
      float val = P;  // mode 1: weight fix40(e); kept for the greedy arg-max as e (not e * inv) below
      u64 w;
      if (mode == 2) {
        float Q = 0.0f;
        for (int i = 0; i < cq; ++i)
          if (sqj[i] == e[h].y) Q = __fmul_rn(cweight(sqz[i], c, rq.mc), rq.inv);
        val = fmaxf(__fsub_rn(P, Q), 0.0f);
        w = fix60(val);
      } else {
        val = cweight(zp, c, rp.mc);
        w = fix40(val);
      }
      e[h].x = rank;  // (reuse: destination slot)
      sw[rank] = w;
      sp[rank] = val;
    }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (lane + 32 * h < cp) sj[e[h].x] = e[h].y;  // token ids in index order (ranks are a permutation: indices are distinct)
  __syncwarp();
  // ---- totals, fallback rule, draw
  u64 w0 = (2 * lane < cp) ? sw[2 * lane] : 0ull, w1 = (2 * lane + 1 < cp) ? sw[2 * lane + 1] : 0ull;
  u64 incl = w0 + w1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  u64 total = __shfl_sync(0xffffffffu, incl, 31);
  const u64 rmin = (job.flags & SPECDEC_RESID_FALLBACK) ? 1152921ull : 0ull;
  bool from_p = (mode == 1);
  if (mode == 2 && total <= rmin) {
    // residual mass (numerically) zero: sample the target row itself (engine/infer_engine.py:319-321)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * lane + h;
      if (i < cp) {
        const float zp = load1<DT>(row_ptr<DT>(rj, r1), sj[i]);
        const float ev = cweight(zp, c, rp.mc);
        sp[i] = ev; sw[i] = fix40(ev);
      }
    }
    __syncwarp();
    w0 = (2 * lane < cp) ? sw[2 * lane] : 0ull; w1 = (2 * lane + 1 < cp) ? sw[2 * lane + 1] : 0ull;
    incl = w0 + w1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    from_p = true;
  }
  long long x = -1;
  if (greedy) {  // largest value, first index on ties (values > 0 only, as the sweep kernels: key == 0 -> token 0)
    float bv = -1.0f; int bi = 0x7FFFFFFF;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * lane + h;
      if (i < cp && sp[i] > bv && (from_p || sp[i] > 0.0f)) { bv = sp[i]; bi = sj[i]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    x = (bi == 0x7FFFFFFF) ? 0ll : (long long)bi;
  } else {
    const u64 target = scale_u24(total, u24_of(us));
    // first index whose inclusive prefix sum exceeds target
    const u64 excl = incl - (w0 + w1);
    int hit = -1;
    if (target < total) {
      if (excl + w0 > target) hit = 2 * lane;
      else if (excl + w0 + w1 > target) hit = 2 * lane + 1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit >= 0);
    if (bal) {
      const int owner = __ffs(bal) - 1;
      hit = __shfl_sync(0xffffffffu, hit, owner);
      x = (long long)sj[hit];
    }
  }
  if (lane == 0) {
    job.next_tok[b] = x;
    if (job.next_prob) {
      float np = 0.0f;
      if (from_p && x >= 0) np = row_prob<DT>(rp, row_ptr<DT>(rj, r1), (int)x, c);
      job.next_prob[b] = np;
    }
    if (job.packed) {
      const long long* toks = job.draft_tokens + (long long)b * g;
      int* pk = job.packed + (long long)b * (g + 2);
      pk[0] = n;
      for (int i = 0; i < g + 1; ++i) pk[1 + i] = -1;
      for (int i = 0; i < n; ++i) pk[1 + i] = (int)toks[i];
      pk[1 + n] = (int)x;
    }
    ws.part_done[b] = CH;  // handled: sample_partial_kernel's CTAs of this sequence return at once
  }
}

template <int DT>
__global__ void __launch_bounds__(SL_WARPS * 32) sample_lists_kernel(DecideJob job, HybridWs ws, int B) {
  __shared__ ListSh lsh[SL_WARPS];
  grid_dependency_wait();
  const int wib = threadIdx.x >> 5;
  const int b = blockIdx.x * SL_WARPS + wib;
  if (b >= B) return;
  sample_lists_sequence<DT>(job, ws, b, lsh[wib]);
}

// plan + draw from the kept-token lists in ONE launch (masked modes with lists: the plan has no exact tasks there, so
// the sequence is decided by plan_sequence and the same warp can draw right away -- one launch and ~8 us less per step)
template <int DT>
__global__ void __launch_bounds__(256) plan_lists_kernel(DecideJob job, HybridWs ws) {
  __shared__ ListSh lsh[8];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  grid_dependency_wait();
  const int wib = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + wib;
  if (b >= (int)(job.rj.R / (job.rj.nT + job.rj.nD))) return;
  plan_sequence<DT>(job, ws, b);
  __syncwarp();
  __threadfence();
  if (__ldcg(&ws.seq_tasks[b]) == 0) sample_lists_sequence<DT>(job, ws, b, lsh[wib]);  // (decided by the plan itself)
}

#include "tail_fused.cuh"
#include "mega.cuh"
