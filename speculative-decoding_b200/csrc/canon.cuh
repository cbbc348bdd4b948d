// canon.cuh -- canonical arithmetic + small device utilities shared by the sm_100a kernels.
//
// The verify path is specified in IEEE-754 fp32 (+,*,fma,/) and integer arithmetic only
// (DESIGN.md section 3), so results are independent of launch geometry and reduction order and
// are reproduced bit for bit by the CPU oracle.  Everything here uses explicit _rn
// intrinsics (never contracted / reassociated by nvcc) and the library is built with
// -fmad=false and without --use_fast_math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace specdec {

typedef unsigned long long u64;

constexpr int DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2;

// 2^f on [-0.5,0.5], degree 5, p(0)=1; max rel. error 1.9e-7 in fp32 Horner/FMA evaluation.
#define SPECDEC_C1 0x1.62e42ap-1f
#define SPECDEC_C2 0x1.ebf9bcp-3f
#define SPECDEC_C3 0x1.c6b752p-5f
#define SPECDEC_C4 0x1.3cea88p-7f
#define SPECDEC_C5 0x1.5bba14p-10f

__device__ __forceinline__ float cexp2(float t) {
  t = fmaxf(t, -125.0f);
  t = fminf(t, 126.0f);
  float r = __fadd_rn(t, 12582912.0f);
  int i = __float_as_int(r) - 0x4B400000;
  float fi = __fsub_rn(r, 12582912.0f);
  float f = __fsub_rn(t, fi);
  float p = SPECDEC_C5;
  p = __fmaf_rn(p, f, SPECDEC_C4);
  p = __fmaf_rn(p, f, SPECDEC_C3);
  p = __fmaf_rn(p, f, SPECDEC_C2);
  p = __fmaf_rn(p, f, SPECDEC_C1);
  p = __fmaf_rn(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (int)((unsigned)i << 23));
}
// cexp2 for arguments known to be <= 126 (every weight: z <= row max => t <= ~0).  Bit-identical to
// cexp2 there; saves the upper clamp and the integer subtraction (0x4B400000 << 23 == 0 mod 2^32).
__device__ __forceinline__ float cexp2_le(float t) {
  t = fmaxf(t, -125.0f);
  const float r = __fadd_rn(t, 12582912.0f);
  const float f = __fsub_rn(t, __fsub_rn(r, 12582912.0f));
  float p = SPECDEC_C5;
  p = __fmaf_rn(p, f, SPECDEC_C4);
  p = __fmaf_rn(p, f, SPECDEC_C3);
  p = __fmaf_rn(p, f, SPECDEC_C2);
  p = __fmaf_rn(p, f, SPECDEC_C1);
  p = __fmaf_rn(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}
// weight of logit z (<= row max m) under (c, mc): exp2(z*c - m*c), one rounding in the exponent argument
__device__ __forceinline__ float cweight(float z, float c, float mc) { return cexp2_le(__fmaf_rn(z, c, -mc)); }
// Two canonical weights per instruction stream: Blackwell's packed fp32x2 FMA/ADD (SASS FFMA2 / FADD2 /
// FMUL2, sm_100+) are IEEE round-to-nearest per component, so the results are bit-identical to two
// cweight() calls while the polynomial costs ~7 instead of 12 issue slots per element.
__device__ __forceinline__ float2 cweight2(float2 z, float2 c, float2 negmc) {
  float2 t = __ffma2_rn(z, c, negmc);
  t.x = fmaxf(t.x, -125.0f);
  t.y = fmaxf(t.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 r = __fadd2_rn(t, magic);
  const float2 fi = __fadd2_rn(r, make_float2(-12582912.0f, -12582912.0f));  // == r - magic
  const float2 f = __ffma2_rn(fi, make_float2(-1.0f, -1.0f), t);              // == t - fi (exact product)
  float2 p = make_float2(SPECDEC_C5, SPECDEC_C5);
  p = __ffma2_rn(p, f, make_float2(SPECDEC_C4, SPECDEC_C4));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_C3, SPECDEC_C3));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_C2, SPECDEC_C2));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_C1, SPECDEC_C1));
  p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
  float2 e;
  e.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23));
  e.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23));
  return e;
}
// cweight2() * 2^40, bit for bit: every Horner step is scaled by a power of two (IEEE rounding commutes with such a
// scaling: no intermediate comes near the subnormal or overflow range -- the polynomial lives in [0.7, 1.5] * 2^40, the
// result in [2^-85, 2^41]), so the fixed-point weight is __float2ull_rz() of the result without the multiply fix40()
// needs, and P * 2^60 = (e * 2^40) * (inv * 2^20) feeds fix60() the same way.  Saves 1.5 issue slots per logit in the
// exact tail (the two elements of a pair are NEIGHBOURS OF ONE ROW here: both halves of the packed registers come out
// of one 16-bit unpack, where pairing the rows cost a MOV per element).
#define SPECDEC_S40(x) ((x) * 1099511627776.0f)
__device__ __forceinline__ float2 cweight2_s40(float2 z, float2 c, float2 negmc) {
  float2 t = __ffma2_rn(z, c, negmc);
  t.x = fmaxf(t.x, -125.0f);
  t.y = fmaxf(t.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 r = __fadd2_rn(t, magic);
  const float2 fi = __fadd2_rn(r, make_float2(-12582912.0f, -12582912.0f));  // == r - magic
  const float2 f = __ffma2_rn(fi, make_float2(-1.0f, -1.0f), t);              // == t - fi (exact product)
  float2 p = make_float2(SPECDEC_S40(SPECDEC_C5), SPECDEC_S40(SPECDEC_C5));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_S40(SPECDEC_C4), SPECDEC_S40(SPECDEC_C4)));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_S40(SPECDEC_C3), SPECDEC_S40(SPECDEC_C3)));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_S40(SPECDEC_C2), SPECDEC_S40(SPECDEC_C2)));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_S40(SPECDEC_C1), SPECDEC_S40(SPECDEC_C1)));
  p = __ffma2_rn(p, f, make_float2(SPECDEC_S40(1.0f), SPECDEC_S40(1.0f)));
  float2 e;
  e.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23));
  e.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23));
  return e;
}
// residuals of two neighbouring tokens from their scaled weights: max(0, P - Q) * 2^60 each, ready for __float2ull_rz
// (ip2 = (inv_p, inv_p) * 2^20, iq2 = (inv_q, inv_q) * 2^20).  The subtraction is SCALAR on purpose: ptxas contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (observed with nvcc 12.9 for sm_100a, -fmad=false notwithstanding), which
// skips the rounding of Q and changed one emitted token in 256 sequences x 128256 tokens; scalar sub.rn is never fused.
__device__ __forceinline__ float2 resid2_s60(float4 e, float2 ip2, float2 iq2) {
  const float2 P = __fmul2_rn(make_float2(e.x, e.y), ip2), Q = __fmul2_rn(make_float2(e.z, e.w), iq2);
  return make_float2(fmaxf(__fsub_rn(P.x, Q.x), 0.0f), fmaxf(__fsub_rn(P.y, Q.y), 0.0f));
}
__device__ __forceinline__ u64 fix40(float x) { return __float2ull_rz(__fmul_rn(x, 1099511627776.0f)); }
__device__ __forceinline__ u64 fix60(float x) { return __float2ull_rz(__fmul_rn(x, 1152921504606846976.0f)); }
__device__ __forceinline__ unsigned u24_of(float u) {
  if (!(u > 0.0f)) return 0u;
  if (u >= 1.0f) return 16777215u;
  return (unsigned)__fmul_rn(u, 16777216.0f);
}
// (S * u24) >> 24 with a 128-bit intermediate
__device__ __forceinline__ u64 scale_u24(u64 S, unsigned u24) {
  u64 hi = __umul64hi(S, (u64)u24), lo = S * (u64)u24;
  return (hi << 40) | (lo >> 24);
}
// (S * q32) >> 32 with a 128-bit intermediate
__device__ __forceinline__ u64 scale_q32(u64 S, u64 q32) {
  u64 hi = __umul64hi(S, q32), lo = S * q32;
  return (hi << 32) | (lo >> 32);
}
// sum of the 8 fixed-point weights of one vector (packed fp32x2 evaluation, bit-identical to scalar)
__device__ __forceinline__ u64 sum_fix40_8(const float (&x)[8], float c, float mc) {
  const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 e = cweight2(make_float2(x[2 * k], x[2 * k + 1]), c2, nmc2);
    s += fix40(e.x) + fix40(e.y);
  }
  return s;
}
// order-preserving float -> uint32 key
__device__ __forceinline__ unsigned fkey(float z) {
  unsigned b = __float_as_uint(__fadd_rn(z, 0.0f));  // -0.0 -> +0.0 so that key order == float order
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
  unsigned b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  return __uint_as_float(b);
}

// MUFU ex2 (SASS MUFU.EX2): used only where ~1e-6 relative accuracy suffices
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- Philox4x32-10, keyed by (seed; offset, global sequence id, lane) ----
__device__ __forceinline__ unsigned philox_word0(u64 seed, u64 offset, unsigned seq, unsigned lane) {
  unsigned c0 = (unsigned)offset, c1 = (unsigned)(offset >> 32), c2 = seq, c3 = lane;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}
__device__ __forceinline__ float philox_uniform(u64 seed, u64 offset, unsigned seq, unsigned lane) {
  return __fmul_rn((float)(philox_word0(seed, offset, seq, lane) >> 8), 0x1p-24f);
}

// ---- 8-element vector access for the three logit dtypes ----
// Vector v covers elements [8v, 8v+8).  `aligned` = row base 16-byte aligned (then every full
// vector is one (bf16/f16) or two (f32) 16-byte loads).  Elements >= V read as -inf (weight 0).
template <int DT> struct Elem;
template <> struct Elem<DT_F32> { typedef float T; };
template <> struct Elem<DT_BF16> { typedef __nv_bfloat16 T; };
template <> struct Elem<DT_F16> { typedef __half T; };

template <int DT>
__device__ __forceinline__ float load1(const void* row, int j) {
  if (DT == DT_F32) return __ldg((const float*)row + j);
  if (DT == DT_BF16) return __uint_as_float(((unsigned)__ldg((const unsigned short*)row + j)) << 16);
  return __half2float(__ushort_as_half(__ldg((const unsigned short*)row + j)));
}

template <int DT>
__device__ __forceinline__ void load8(const void* row, int v, int V, bool aligned, float (&x)[8]) {
  const int j0 = v * 8;
  if (aligned && j0 + 8 <= V) {
    if (DT == DT_F32) {
      const float4* p = (const float4*)((const float*)row + j0);
      float4 a = __ldg(p), b = __ldg(p + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
      uint4 a = __ldg((const uint4*)((const unsigned short*)row + j0));
      unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (DT == DT_BF16) {
          x[2 * k] = __uint_as_float(w[k] << 16);
          x[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
        } else {
          __half2 h = *reinterpret_cast<__half2*>(&w[k]);
          float2 f = __half22float2(h);
          x[2 * k] = f.x; x[2 * k + 1] = f.y;
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (j0 + k < V) ? load1<DT>(row, j0 + k) : -INFINITY;
  }
}

// Raw (still packed) 8-element vector: lets a thread keep many loads in flight without paying the
// registers of the unpacked floats.  bf16/f16: 4 registers; f32: 8.
template <int DT> struct Raw8 { uint4 a; };
template <> struct Raw8<DT_F32> { uint4 a, b; };

template <int DT>
__device__ __forceinline__ Raw8<DT> load_raw8(const void* row, int v, int V, bool aligned) {
  Raw8<DT> r;
  const int j0 = v * 8;
  if (aligned && j0 + 8 <= V) {
    if constexpr (DT == DT_F32) {
      const uint4* p = (const uint4*)((const float*)row + j0);
      r.a = __ldg(p); r.b = __ldg(p + 1);
    } else {
      r.a = __ldg((const uint4*)((const unsigned short*)row + j0));
    }
  } else {  // ragged tail / unaligned row: scalar loads, -inf padding
    if constexpr (DT == DT_F32) {
      unsigned w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = (j0 + k < V) ? __float_as_uint(__ldg((const float*)row + j0 + k)) : 0xFF800000u;
      r.a = make_uint4(w[0], w[1], w[2], w[3]); r.b = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      const unsigned ninf = (DT == DT_BF16) ? 0xFF80u : 0xFC00u;
      unsigned w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned lo = (j0 + 2 * k < V) ? (unsigned)__ldg((const unsigned short*)row + j0 + 2 * k) : ninf;
        const unsigned hi = (j0 + 2 * k + 1 < V) ? (unsigned)__ldg((const unsigned short*)row + j0 + 2 * k + 1) : ninf;
        w[k] = lo | (hi << 16);
      }
      r.a = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  return r;
}
template <int DT>
__device__ __forceinline__ void unpack8(const Raw8<DT>& r, float (&x)[8]) {
  if constexpr (DT == DT_F32) {
    x[0] = __uint_as_float(r.a.x); x[1] = __uint_as_float(r.a.y); x[2] = __uint_as_float(r.a.z); x[3] = __uint_as_float(r.a.w);
    x[4] = __uint_as_float(r.b.x); x[5] = __uint_as_float(r.b.y); x[6] = __uint_as_float(r.b.z); x[7] = __uint_as_float(r.b.w);
  } else {
    const unsigned w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if constexpr (DT == DT_BF16) {
        x[2 * k] = __uint_as_float(w[k] << 16);
        x[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
        x[2 * k] = f.x; x[2 * k + 1] = f.y;
      }
    }
  }
}

// ---- the CTA's compute group ----
// Ordinary kernels: all threads of the CTA, __syncthreads().  CTAs of exactly 288 threads are warp-specialised
// (rowfast_tma_kernel, verify_mega_kernel): warps 0-7 form the compute group (named barrier 1, 256 threads), warp 8 is
// a service warp (TMA producer / planner) that never calls the block-wide helpers below.
constexpr int WS_CTA_THREADS = 288, WS_COMPUTE_THREADS = 256;
__device__ __forceinline__ int cta_nthreads() { return blockDim.x == WS_CTA_THREADS ? WS_COMPUTE_THREADS : (int)blockDim.x; }
__device__ __forceinline__ void cta_sync() {
  if (blockDim.x == WS_CTA_THREADS) asm volatile("bar.sync 1, %0;" ::"n"(WS_COMPUTE_THREADS) : "memory");
  else __syncthreads();
}

// ---- block-wide reductions (up to 32 warps); `sh` = 33-entry scratch ----
// Exact (mod 2^64) warp sum of u64 values: three independent hardware integer reductions (REDUX) over 22 / 22 / 20-bit
// pieces instead of five dependent 64-bit shuffle-add rounds -- a third of the latency, half the instructions.
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
  const unsigned a = (unsigned)v & 0x3FFFFFu, b = (unsigned)(v >> 22) & 0x3FFFFFu, c = (unsigned)(v >> 44);
  const unsigned sa = __reduce_add_sync(0xffffffffu, a), sb = __reduce_add_sync(0xffffffffu, b),
                 sc = __reduce_add_sync(0xffffffffu, c);
  return (u64)sa + ((u64)sb << 22) + ((u64)sc << 44);
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// All threads get the result.  Two barriers; safe to call back to back with the same scratch.
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64* sh) {
  v = warp_sum_u64(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (cta_nthreads() + 31) >> 5;
  if (lane == 0) sh[w] = v;
  cta_sync();
  u64 t = (lane < nw) ? sh[lane] : 0ull;
  t = warp_sum_u64(t);
  cta_sync();
  return t;
}
__device__ __forceinline__ float block_max_f(float v, float* sh) {
  v = warp_max_f(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (cta_nthreads() + 31) >> 5;
  if (lane == 0) sh[w] = v;
  cta_sync();
  float t = (lane < nw) ? sh[lane] : -INFINITY;
  t = warp_max_f(t);
  cta_sync();
  return t;
}
__device__ __forceinline__ float block_min_f(float v, float* sh) {
  v = warp_min_f(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (cta_nthreads() + 31) >> 5;
  if (lane == 0) sh[w] = v;
  cta_sync();
  float t = (lane < nw) ? sh[lane] : INFINITY;
  t = warp_min_f(t);
  cta_sync();
  return t;
}
__device__ __forceinline__ unsigned block_min_u32(unsigned v, unsigned* sh) {
  v = __reduce_min_sync(0xffffffffu, v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (cta_nthreads() + 31) >> 5;
  if (lane == 0) sh[w] = v;
  cta_sync();
  unsigned t = (lane < nw) ? sh[lane] : 0xFFFFFFFFu;
  t = __reduce_min_sync(0xffffffffu, t);
  cta_sync();
  return t;
}

}  // namespace specdec
