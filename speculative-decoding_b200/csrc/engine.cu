// engine.cu -- the two small device-side pieces of the decode LOOPS around the verify step (sm_100a).
//
//  * topk_ids_kernel: ids of the k largest logits of every row (value descending, index ascending on ties) -- the
//    "filler" tokens the n-gram-assisted loop feeds back into its table after every accepted position
//    (ngram_assisted/ngram_assisted.py:149-155: `torch.topk(p[..., i, :], filler_top_k)`; the probabilities are a
//    monotone function of the logits, so the ids are taken from the logits and no V-wide probability row is ever
//    written).  One CTA per row, one pass: 16-byte loads, a per-thread sorted top-KT list in registers that is
//    touched only when a vector's maximum beats the thread's current k-th value, then k rounds of a block arg-max
//    over the list heads.  HBM-bound: V * elem bytes per row.
//  * batch_writeback_kernel: the ragged per-sequence bookkeeping after a batched verify
//    (engine/infer_engine.py:300-336: accepted count, corrected token at the first rejection, zeroed tail,
//    finished flags) as ONE launch on device-resident state, so the batched loop runs without a host read-back
//    per step and its step is capturable into a CUDA graph.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include "../../include/specdec_b200.h"

namespace specdec {

constexpr int TK_T = 256;    // threads per row
constexpr int TK_MAX = 8;    // ids per pass (register list length); larger k = more passes

template <int DT>
__device__ __forceinline__ void tk_load8(const void* row, int j0, int V, bool aligned, float (&x)[8]) {
  if (aligned && j0 + 8 <= V) {
    if (DT == 0) {
      const float4* p = (const float4*)((const float*)row + j0);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
      const uint4 a = __ldg((const uint4*)((const unsigned short*)row + j0));
      const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (DT == 1) {
          x[2 * k] = __uint_as_float(w[k] << 16);
          x[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
        } else {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
          x[2 * k] = f.x; x[2 * k + 1] = f.y;
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = j0 + k;
      float v = -INFINITY;
      if (j < V) {
        if (DT == 0) v = __ldg((const float*)row + j);
        else if (DT == 1) v = __uint_as_float(((unsigned)__ldg((const unsigned short*)row + j)) << 16);
        else v = __half2float(__ushort_as_half(__ldg((const unsigned short*)row + j)));
      }
      x[k] = v;
    }
  }
}

// (value, index) order: larger value first, smaller index first among equal values; NaN never selected
__device__ __forceinline__ bool tk_before(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

template <int DT>
__global__ void __launch_bounds__(TK_T) topk_ids_kernel(const void* logits, long long stride, int V, int k, long long* out) {
  __shared__ float s_v[TK_T / 32];
  __shared__ int s_i[TK_T / 32];
  __shared__ int s_t[TK_T / 32];
  __shared__ float s_wv;
  __shared__ int s_wi, s_wt;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const size_t es = (DT == 0) ? 4 : 2;
  const void* row = (const char*)logits + (size_t)blockIdx.x * (size_t)stride * es;
  const bool aligned = (((size_t)row) & 15) == 0;
  long long* dst = out + (long long)blockIdx.x * k;
  // elements strictly after (cv, ci) in the order are eligible in this pass (first pass: everything)
  float cv = INFINITY;
  int ci = -1;
  for (int done = 0; done < k; done += TK_MAX) {
    const int want = min(TK_MAX, k - done);
    float lv[TK_MAX];
    int li[TK_MAX];
#pragma unroll
    for (int q = 0; q < TK_MAX; ++q) { lv[q] = -INFINITY; li[q] = 0x7FFFFFFF; }
    for (int j0 = tid * 8; j0 < V; j0 += TK_T * 8) {
      float x[8];
      tk_load8<DT>(row, j0, V, aligned, x);
      const float vm = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
      if (vm >= lv[TK_MAX - 1]) {  // rare after the first few vectors
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v = x[e];
          const int j = j0 + e;
          if (j < V && (done == 0 || tk_before(cv, ci, v, j)) && tk_before(v, j, lv[TK_MAX - 1], li[TK_MAX - 1])) {
            lv[TK_MAX - 1] = v; li[TK_MAX - 1] = j;
#pragma unroll
            for (int q = TK_MAX - 1; q > 0; --q) {  // one bubble pass keeps the list sorted
              if (tk_before(lv[q], li[q], lv[q - 1], li[q - 1])) {
                const float tv = lv[q]; lv[q] = lv[q - 1]; lv[q - 1] = tv;
                const int ti = li[q]; li[q] = li[q - 1]; li[q - 1] = ti;
              }
            }
          }
        }
      }
    }
    // `want` rounds: block arg-max over the list heads, the winner pops its head
    for (int r = 0; r < want; ++r) {
      float v = lv[0];
      int i = li[0], t = tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o), ot = __shfl_xor_sync(0xffffffffu, t, o);
        if (tk_before(ov, oi, v, i)) { v = ov; i = oi; t = ot; }
      }
      if (lane == 0) { s_v[w] = v; s_i[w] = i; s_t[w] = t; }
      __syncthreads();
      if (w == 0) {
        v = (lane < TK_T / 32) ? s_v[lane] : -INFINITY;
        i = (lane < TK_T / 32) ? s_i[lane] : 0x7FFFFFFF;
        t = (lane < TK_T / 32) ? s_t[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, i, o), ot = __shfl_xor_sync(0xffffffffu, t, o);
          if (tk_before(ov, oi, v, i)) { v = ov; i = oi; t = ot; }
        }
        if (lane == 0) { s_wv = v; s_wi = i; s_wt = t; }
      }
      __syncthreads();
      const int wi = s_wi, wt = s_wt;
      const float wv = s_wv;
      if (tid == 0) dst[done + r] = (wi == 0x7FFFFFFF) ? -1ll : (long long)wi;  // fewer than k selectable elements
      if (tid == wt) {
#pragma unroll
        for (int q = 0; q < TK_MAX - 1; ++q) { lv[q] = lv[q + 1]; li[q] = li[q + 1]; }
        lv[TK_MAX - 1] = -INFINITY; li[TK_MAX - 1] = 0x7FFFFFFF;
      }
      cv = wv; ci = wi;
      __syncthreads();
    }
  }
}

// one thread per sequence; generated[b, step + j] for j < g is rewritten exactly as engine/infer_engine.py:300-336 does
__global__ void batch_writeback_kernel(int B, int g, const int* n_accepted, const int* first_stop, const long long* next_token,
                                       long long* generated, long long gen_stride, const long long* step_dev, long long step,
                                       unsigned char* finished, long long* n_acc, const long long* end_tokens, int n_end,
                                       int* n_active_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (finished[b]) return;
  const long long s0 = step_dev ? *step_dev : step;
  const int n = n_accepted[b], fs = first_stop[b];
  const long long x = next_token[b];
  const bool hit_end = fs >= 0;
  const int acc = hit_end ? fs + 1 : n;     // accepted drafts end at the first accepted end token (:310-312)
  const bool rejected = !hit_end && n < g;
  n_acc[b] += acc;
  long long* cur = generated + (long long)b * gen_stride + s0;
  if (rejected) cur[n] = x;                 // corrected token at the first rejection (:326)
  for (int j = acc + 1; j < g; ++j) cur[j] = 0;  // zeros after it (:333-336)
  bool x_is_end = false;
  for (int e = 0; e < n_end; ++e) x_is_end |= (end_tokens[e] == x);
  if (hit_end || (rejected && x_is_end)) finished[b] = 1;
  else if (n_active_out) atomicAdd(n_active_out, 1);
}

// ---- all-gather of the packed per-sequence results by peer-to-peer stores over NVLink / NVSwitch ----
// One CTA per peer copies this rank's block of packed int32 words into the peer's gather buffer (plain stores into
// peer-mapped memory), makes them visible system-wide and then releases the peer's arrival flag of this rank with
// the step's sequence number.  No NCCL kernel, no host round trip: the launch rides the verify stream right behind
// the step that produced the words.
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) peer_publish_kernel(const int* src, int n_words, int* const* peer_bufs, long long dst_off_words,
                                                           int* const* peer_flags, long long flag_off_words, int seq, int world,
                                                           const int* wait_flags, int wait_seq, int* status) {
  if ((int)blockIdx.x == world) {
    // optional extra CTA: the reader side of an EARLIER step (one launch instead of two per step) -- lane r waits until
    // rank r's flag of that step's slot has reached wait_seq
    if (threadIdx.x < 32) {
      bool ok = true;
      if ((int)threadIdx.x < world) {
        ok = false;
        for (unsigned it = 0; it < (1u << 24); ++it) {
          int v;
          asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(wait_flags + threadIdx.x) : "memory");
          if (v - wait_seq >= 0) { ok = true; break; }
          __nanosleep(100);
        }
      }
      if (!__all_sync(0xffffffffu, ok) && threadIdx.x == 0 && status) *status = 1;
    }
    return;
  }
  int* dst = peer_bufs[blockIdx.x] + dst_off_words;
  if ((((size_t)src | (size_t)dst) & 15) == 0) {
    const int n4 = n_words >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(src)[i];
    for (int i = (n4 << 2) + threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(peer_flags[blockIdx.x] + flag_off_words, seq);
}
// one warp: lane r waits until rank r's flag of the slot has reached seq (bounded: ~2 s, then *status = 1)
__global__ void peer_wait_kernel(const int* flags, int world, int seq, int* status) {
  const int lane = threadIdx.x;
  bool ok = true;
  if (lane < world) {
    ok = false;
    for (unsigned it = 0; it < (1u << 24); ++it) {
      if (ld_acquire_sys(flags + lane) - seq >= 0) { ok = true; break; }
      __nanosleep(100);
    }
  }
  if (!__all_sync(0xffffffffu, ok) && lane == 0 && status) *status = 1;
}

}  // namespace specdec

extern "C" int specdec_peer_publish(const int32_t* packed_local, int n_words, void* const* peer_bufs_dev, int64_t dst_off_words,
                                    int world, void* const* peer_flags_dev, int64_t flag_off_words, int32_t seq,
                                    const int32_t* wait_flags_local, int32_t wait_seq, int32_t* status, specdec_stream_t stream) {
  if (!packed_local || n_words <= 0 || !peer_bufs_dev || !peer_flags_dev || world <= 0 || world > 32) return SPECDEC_ERR_ARG;
  specdec::peer_publish_kernel<<<world + (wait_flags_local ? 1 : 0), 256, 0, (cudaStream_t)stream>>>(
      (const int*)packed_local, n_words, (int* const*)peer_bufs_dev, dst_off_words, (int* const*)peer_flags_dev, flag_off_words, seq,
      world, (const int*)wait_flags_local, wait_seq, (int*)status);
  return (int)cudaGetLastError();
}
extern "C" int specdec_peer_wait(const int32_t* flags_local, int world, int32_t seq, int32_t* status, specdec_stream_t stream) {
  if (!flags_local || world <= 0 || world > 32) return SPECDEC_ERR_ARG;
  specdec::peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const int*)flags_local, world, seq, (int*)status);
  return (int)cudaGetLastError();
}

extern "C" int specdec_topk_ids(const void* logits, int dtype, int64_t rows, int V, int64_t row_stride, int k, int64_t* out_ids,
                                specdec_stream_t stream) {
  if (rows < 0 || V <= 0 || k <= 0 || k > V || !out_ids || (rows > 0 && !logits)) return SPECDEC_ERR_ARG;
  if (rows == 0) return 0;
  if (rows > 0x7FFFFFFF) return SPECDEC_ERR_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case SPECDEC_F32: specdec::topk_ids_kernel<0><<<(unsigned)rows, specdec::TK_T, 0, st>>>(logits, row_stride, V, k, (long long*)out_ids); break;
    case SPECDEC_BF16: specdec::topk_ids_kernel<1><<<(unsigned)rows, specdec::TK_T, 0, st>>>(logits, row_stride, V, k, (long long*)out_ids); break;
    case SPECDEC_F16: specdec::topk_ids_kernel<2><<<(unsigned)rows, specdec::TK_T, 0, st>>>(logits, row_stride, V, k, (long long*)out_ids); break;
    default: return SPECDEC_ERR_DTYPE;
  }
  return (int)cudaGetLastError();
}

extern "C" int specdec_batch_writeback(int B, int gamma, const int32_t* n_accepted, const int32_t* first_stop,
                                       const int64_t* next_token, int64_t* generated, int64_t gen_stride,
                                       const int64_t* step_dev, int64_t step, uint8_t* finished, int64_t* n_acc,
                                       const int64_t* end_tokens, int n_end, int32_t* n_active_out,
                                       specdec_stream_t stream) {
  if (B < 0 || gamma <= 0 || !n_accepted || !first_stop || !next_token || !generated || !finished || !n_acc) return SPECDEC_ERR_ARG;
  if (n_end > 0 && !end_tokens) return SPECDEC_ERR_ARG;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_active_out) {
    cudaError_t e = cudaMemsetAsync(n_active_out, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  specdec::batch_writeback_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, gamma, n_accepted, first_stop, (const long long*)next_token,
                                                                  (long long*)generated, gen_stride, (const long long*)step_dev, step,
                                                                  finished, (long long*)n_acc, (const long long*)end_tokens, n_end,
                                                                  n_active_out);
  return (int)cudaGetLastError();
}
