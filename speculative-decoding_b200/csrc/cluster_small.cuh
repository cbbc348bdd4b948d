// cluster_small.cuh -- small batches: the whole verify step of the plain modes as ONE launch, one thread-block CLUSTER
// per sequence (included by verify.cu after hybrid.cuh, inside namespace specdec).
//
// At B <= ~64 the three-launch pipeline is bound by its own latency chain, not by bytes: memset -> row kernel (one CTA
// streams a whole 256 KB row) -> plan -> fused tail, ~50 us of device time at B = 1 for 2.3 MB of logits
// (profiles/r2_small_launches.csv), and sampling/speculative_decoding.py *is* batch 1.  Here the CL CTAs of a cluster
// (co-scheduled by the hardware: the waits between them cannot dead-lock, unlike ticket-ordered independent CTAs)
// split every row of ONE sequence CL ways:
//   1. row statistics: CTA `ch` sweeps slice ch of all 2*gamma+1 rows (several rows' loads in flight per thread: the
//      phase is latency bound), online (max, sum of MUFU ex2), partials to the workspace;        barrier.cluster
//   2. warp 0 of CTA 0 merges the partials into the RowOut records and runs plan_sequence (accept tests with the
//      1e-3 margin, hybrid.cuh) -- the same code the pipeline's plan_kernel runs;                  barrier.cluster
//   3. every CTA runs mega_item (mega.cuh: tail_item with the exchange through self-validating words) on its slice of
//      the deciding row pair: canonical weights cached in shared memory, exact normalisers, residual partial sums,
//      token location.
// MEASURED (B200, graph-replayed step, bf16, gamma 4, V 128256): 36 / 68 / 121 us at B = 1 / 32 / 64 against 34 / 47 / 67 us
// for the three-launch pipeline -- the phases of one sequence run back to back here (statistics, plan by ONE warp
// while 127 wait, exact item), where the pipeline overlaps them across sequences.  The path is therefore OPT-IN
// (specdec_set_option("small_b", B_max)); it stays built and parity-tested because it is the only path whose
// inter-CTA waits are co-scheduled by hardware (no ticket-order argument), e.g. under MPS SM limits.
// Same integers as the pipeline => bit-identical results.  Exchanges go through the (L2-resident) workspace with
// release/acquire cluster barriers; nothing spins except tail_item's own bounded group waits.
#pragma once

constexpr int CS_T = TF_T;  // threads per CTA (tail_item's work loops assume TF_T)
constexpr int CS_RG = 5;    // rows swept together in phase 1 (their loads are in flight together)

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float block_sum_f(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (cta_nthreads() + 31) >> 5;
  if (lane == 0) sh[w] = v;
  cta_sync();
  float t = (lane < nw) ? sh[lane] : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  cta_sync();
  return t;
}

template <int DT, bool GREEDY>
__global__ void __launch_bounds__(CS_T) verify_cluster_kernel(DecideJob job, HybridWs ws, float2* cpart, int segs_per_cta, int CL) {
  extern __shared__ __align__(16) float4 ecache[];
  __shared__ TailSh sh;
  const RowJob& rj = job.rj;
  const int b = blockIdx.x / CL, ch = blockIdx.x - b * CL;  // cluster dims (CL,1,1): the CTAs of a cluster are consecutive
  const int tid = threadIdx.x, lane = tid & 31;
  const int V = rj.V, NV = (V + 7) >> 3, rps = rj.nT + rj.nD;
  const float c = rj.c;
  // ---- 1. statistics of my slice of every row of the sequence
  const int per = (NV + CL - 1) / CL, v0 = min(NV, ch * per), v1 = min(NV, v0 + per);
  for (int k0 = 0; k0 < rps; k0 += CS_RG) {
    float m[CS_RG], s[CS_RG];
    const void* rowp[CS_RG];
    bool al[CS_RG];
#pragma unroll
    for (int q = 0; q < CS_RG; ++q) {
      m[q] = -INFINITY; s[q] = 0.0f;
      rowp[q] = seq_row_ptr<DT>(rj, b, min(k0 + q, rps - 1));
      al[q] = (((size_t)rowp[q]) & 15) == 0;
    }
    for (int v = v0 + tid; v < v1; v += 2 * CS_T) {
      Raw8<DT> ra[CS_RG], rb[CS_RG];
      const bool two = v + CS_T < v1;
#pragma unroll
      for (int q = 0; q < CS_RG; ++q) {
        if (k0 + q < rps) {
          ra[q] = load_raw8<DT>(rowp[q], v, V, al[q]);
          rb[q] = load_raw8<DT>(rowp[q], two ? v + CS_T : v, V, al[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < CS_RG; ++q) {
        if (k0 + q < rps) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) continue;
            float x[8];
            unpack8<DT>(h ? rb[q] : ra[q], x);
            const float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
            if (vm > m[q]) {
              s[q] = __fmul_rn(s[q], ex2_approx(__fmul_rn(__fsub_rn(m[q], vm), c)));
              m[q] = vm;
            }
            const float mc = (m[q] > -INFINITY) ? __fmul_rn(m[q], c) : 0.0f;  // (only -inf so far: -inf*c + inf would be NaN)
#pragma unroll
            for (int e = 0; e < 8; ++e) s[q] = __fadd_rn(s[q], ex2_approx(__fmaf_rn(x[e], c, -mc)));
          }
        }
      }
    }
    // per-warp (max, sum) of every row of the group -> shared memory; one barrier for the whole group
    float* wm = reinterpret_cast<float*>(ecache);  // [CS_T / 32][CS_RG][2] floats of the (still unused) weight cache
#pragma unroll
    for (int q = 0; q < CS_RG; ++q) {
      const float M = warp_max_f(m[q]);
      float sv = (m[q] > -INFINITY) ? __fmul_rn(s[q], ex2_approx(__fmul_rn(__fsub_rn(m[q], M), c))) : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (lane == 0) { wm[((tid >> 5) * CS_RG + q) * 2] = M; wm[((tid >> 5) * CS_RG + q) * 2 + 1] = sv; }
    }
    __syncthreads();
    if (tid < CS_RG && k0 + tid < rps) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < CS_T / 32; ++w) M = fmaxf(M, wm[(w * CS_RG + tid) * 2]);
      float S = 0.0f;
#pragma unroll
      for (int w = 0; w < CS_T / 32; ++w) {
        const float mw = wm[(w * CS_RG + tid) * 2];
        if (mw > -INFINITY) S = __fadd_rn(S, __fmul_rn(wm[(w * CS_RG + tid) * 2 + 1], ex2_approx(__fmul_rn(__fsub_rn(mw, M), c))));
      }
      cpart[((size_t)b * rps + k0 + tid) * CL + ch] = make_float2(M, S);
    }
    __syncthreads();
  }
  cluster_sync_all();
  // ---- 2. merge + plan: one warp of the cluster
  if (ch == 0 && tid < 32) {
    for (int k = lane; k < rps; k += 32) {
      const float2* pp = cpart + ((size_t)b * rps + k) * CL;
      float M = -INFINITY;
      for (int i = 0; i < CL; ++i) M = fmaxf(M, __ldcg(&pp[i]).x);
      float S = 0.0f;
      for (int i = 0; i < CL; ++i) {
        const float2 t = __ldcg(&pp[i]);
        if (t.x > -INFINITY) S = __fadd_rn(S, __fmul_rn(t.y, ex2_approx(__fmul_rn(__fsub_rn(t.x, M), c))));
      }
      RowOut o;
      o.m = M; o.mc = __fmul_rn(M, c); o.inv = __fdiv_rn(1.0f, S);
      o.cut = -INFINITY; o.jcut = V; o.flags = 0; o.Sfix = 0;
      rj.out[(size_t)b * rps + k] = o;
    }
    __syncwarp();
    __threadfence();
    plan_sequence<DT>(job, ws, b);
    __syncwarp();
    __threadfence();
  }
  cluster_sync_all();
  // ---- 3. everything after the plan, slice ch of CL
  __shared__ volatile int credit;
  const int4 rec = __ldcg((const int4*)(ws.samp + b * SAMP_N)), rec2 = __ldcg((const int4*)(ws.samp + b * SAMP_N) + 1);
  mega_item<DT, GREEDY, 4>(job, ws, b, ch, CL, segs_per_cta, ecache, sh, rec2.y, rec, rec2.x, &credit);
}
