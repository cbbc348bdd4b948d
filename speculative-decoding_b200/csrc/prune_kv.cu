// prune_kv.cu -- per-sequence KV-cache rollback on a static cache (sm_100a).
// Replaces prune_cache / prune_tuple_cache (utils/caching.py:6-55): the reference drops the last n
// positions of every layer tensor [B,H,S,D] as a zero-copy view, which only works for one uniform n.
// Batched speculative decoding needs a different n per sequence (drafter: gamma-n_b, target:
// gamma-n_b+1, sampling/speculative_decoding.py:163-165), so here the cache is static
// [B,H,S_max,D] with a length vector; rollback = length update (one tiny launch) and, optionally, a zero fill of the
// discarded tail (16-byte stores, one CTA per (tensor, sequence)).  The valid prefix equals the reference's view.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/specdec_b200.h"

namespace specdec {

// one CTA per (tensor, sequence): the discarded positions of all H heads, 16-byte stores.  The zero fill is optional
// (the length vector alone defines the valid prefix); with it the cache bytes equal the reference's view + zeros.
__global__ void __launch_bounds__(128) prune_fill_kernel(void* const* tensors, int H, long long S_max, long long D,
                                                         int eb, const int* seq_lens, const int* discard) {
  const int t = blockIdx.y, b = blockIdx.x;
  const int old = seq_lens[b];
  int d = discard[b];
  d = d < 0 ? 0 : (d > old ? old : d);
  if (d == 0) return;
  const int nw = old - d;
  const size_t row_bytes = (size_t)D * eb, nbytes = (size_t)d * row_bytes, head_stride = (size_t)S_max * row_bytes;
  char* base = (char*)tensors[t] + ((size_t)b * H * S_max + nw) * row_bytes;  // head 0
  if (((((size_t)base) | nbytes | head_stride) & 15) == 0) {
    const unsigned n16 = (unsigned)(nbytes >> 4), total = n16 * (unsigned)H;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (unsigned i = threadIdx.x; i < total; i += blockDim.x) {
      const unsigned h = i / n16, k = i - h * n16;
      reinterpret_cast<uint4*>(base + (size_t)h * head_stride)[k] = z;
    }
  } else {
    for (int h = 0; h < H; ++h)
      for (size_t i = threadIdx.x; i < nbytes; i += blockDim.x) base[(size_t)h * head_stride + i] = 0;
  }
}
__global__ void prune_lens_kernel(int B, int* seq_lens, const int* discard) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int old = seq_lens[b];
  int d = discard[b];
  d = d < 0 ? 0 : (d > old ? old : d);
  seq_lens[b] = old - d;
}

}  // namespace specdec

extern "C" int specdec_prune_kv(void* const* tensor_ptrs, int n_tensors, int B, int H, int64_t S_max, int64_t D,
                                int elem_bytes, int32_t* seq_lens, const int32_t* discard, int zero_fill,
                                specdec_stream_t stream) {
  if (B < 0 || H <= 0 || S_max <= 0 || D <= 0 || elem_bytes <= 0 || n_tensors < 0) return SPECDEC_ERR_ARG;
  if (!seq_lens || !discard || (n_tensors > 0 && !tensor_ptrs)) return SPECDEC_ERR_ARG;
  if (B == 0) return 0;
  if (B > 65535 || n_tensors > 65535) return SPECDEC_ERR_RANGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_fill && n_tensors > 0) {
    dim3 grid((unsigned)B, (unsigned)n_tensors);
    specdec::prune_fill_kernel<<<grid, 128, 0, st>>>(tensor_ptrs, H, S_max, D, elem_bytes, seq_lens, discard);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  specdec::prune_lens_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, seq_lens, discard);
  return (int)cudaGetLastError();
}
