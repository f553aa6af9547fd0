// Speaker (Att2in2) kernels other than the dense contractions: additive attention forward /
// backward, LSTM (maxout) pointwise forward / backward, sampling + next-input gather, the
// straight-through and log-prob logit gradients, and small reductions.
//
// HBM-bound kernels: bf16 storage for the region tensors (att_e, p_att), 16-byte vector loads,
// one CTA per batch row for attention (warp per region, online softmax), warp-shuffle reductions.
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace coopcap {

// RNG site ids (Philox "stream" = site base + step)
enum : uint64_t {
  SITE_DROP_ATT = 1ull << 32,
  SITE_DROP_EMBED = 2ull << 32,
  SITE_DROP_CORE = 3ull << 32,
  SITE_NOISE = 4ull << 32,
  SITE_PARTIAL = 5ull << 32,    // partial-sampling row selection
  SITE_SCHED = 6ull << 32,      // scheduled-sampling row selection
};

constexpr int NOISE_ROUNDS = 7;   // Philox rounds of the per-logit noise site (see common.cuh)

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 float8_to_bf16x8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo);
  o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = o;
}

// 4 dropout keep decisions for elements [4*g, 4*g+3] of a site
__device__ __forceinline__ void keep4(const uint8_t* inj, int64_t idx4, uint64_t seed,
                                      uint64_t stream, float p, bool (&k)[4]) {
  if (inj) {
    const uchar4 b = *reinterpret_cast<const uchar4*>(inj + idx4 * 4);
    k[0] = b.x; k[1] = b.y; k[2] = b.z; k[3] = b.w;
  } else {
    const uint4 r = Philox::gen(seed, stream, static_cast<uint64_t>(idx4));
    k[0] = Philox::u01(r.x) >= p; k[1] = Philox::u01(r.y) >= p;
    k[2] = Philox::u01(r.z) >= p; k[3] = Philox::u01(r.w) >= p;
  }
}

// ---------------------------------------------------------------------------------------------
// noise transforms shared by the sampler (forward) and the straight-through backward, which must
// rebuild bit-identical scores.  `fast` selects hardware approximations (Philox noise: production)
// versus libm-accurate logs (injected noise: parity with the reference's torch.log).
// ---------------------------------------------------------------------------------------------
// single-instruction base-2 transcendentals (the libm-style wrappers add range / denormal
// handling worth ~6 instructions per call, which made the sampler issue-bound)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// -log(u) for u in (0, 1]; lg2.approx has a 2^-22 absolute error, useless for u -> 1, where the
// log1p series takes over (branch-free select)
__device__ __forceinline__ float neg_log_fast(float u) {
  const float d = 1.f - u;
  const float series = d * fmaf(d, fmaf(d, 0.33333334f, 0.5f), 1.f);   // d < 1/16: rel. error < d^3/4
  const float direct = -0.69314718f * lg2_ftz(u);
  return d < 0.0625f ? series : direct;
}
__device__ __forceinline__ float gumbel_of(float u, bool fast) {
  if (fast) {
    const float e = neg_log_fast(fmaxf(u, 1e-20f));          // Exp(1)
    return -0.69314718f * lg2_ftz(e);
  }
  return -logf(-logf(u + 1e-20f) + 1e-20f);                  // gumbel.py:6-11
}
// -log(E) for the multinomial race; E injected (precise) or E = -log(1-u) from a Philox uniform
__device__ __forceinline__ float neg_log_exp1(float n, bool injected) {
  if (injected) return -logf(n);
  const float e = neg_log_fast(fmaxf(1.f - n, 1e-20f));
  return -0.69314718f * lg2_ftz(e);
}
// 4 noise values for elements [4*v4, 4*v4+3] of a row: injected (fp32 row) or Philox uniforms
__device__ __forceinline__ void noise4(const float* inj_row, int v4, uint64_t seed, uint64_t stream,
                                       uint64_t ctr, float (&u)[4]) {
  if (inj_row) {
    const float4 t = *reinterpret_cast<const float4*>(inj_row + 4 * v4);
    u[0] = t.x; u[1] = t.y; u[2] = t.z; u[3] = t.w;
  } else {
    const uint4 r = Philox::gen_r<NOISE_ROUNDS>(seed, stream, ctr);
    u[0] = Philox::u01(r.x); u[1] = Philox::u01(r.y); u[2] = Philox::u01(r.z); u[3] = Philox::u01(r.w);
  }
}
// score whose softmax is y (straight-through modes) for one logit
__device__ __forceinline__ float st_score(int mode, float x, float u, float inv_tau, bool fast) {
  return (mode == 2 /*ST_GUMBEL*/ || mode == 5 /*PS_GUMBEL*/) ? (x + gumbel_of(u, fast)) * inv_tau
                                                              : x * inv_tau;
}

// ------------------------------------------------------------------------------------------
// embedding of the fed token: x = dropout(relu(embed[tok]))            (AttModel.py:74-76)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void embed_row(const float* __restrict__ embed, int64_t tok, int E,
                                          const uint8_t* keep_row, uint64_t seed, uint64_t stream,
                                          int64_t elem_base, float drop_p, __nv_bfloat16* __restrict__ dst) {
  const float sc = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float4* src = reinterpret_cast<const float4*>(embed + tok * E);
  for (int i = threadIdx.x; i < E / 4; i += blockDim.x) {
    float4 v = __ldg(src + i);
    float x[4] = {fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f)};
    if (drop_p > 0.f) {
      bool k[4];
      keep4(keep_row, keep_row ? i : (elem_base >> 2) + i, seed, stream, drop_p, k);
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = k[j] ? x[j] * sc : 0.f;
    }
    __nv_bfloat162 a = __floats2bfloat162_rn(x[0], x[1]), b2 = __floats2bfloat162_rn(x[2], x[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b2);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
}

// fp16 logits (z16_all): 8 / 4 consecutive values as floats
__device__ __forceinline__ void f16x8_to_float(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void f16x4_to_float(const __half* p, float (&f)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

// block-wide reductions for 256-thread CTAs (8 warps)
__device__ __forceinline__ float block_sum_256(float v, float* sm /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += sm[i];
  return t;
}
__device__ __forceinline__ float block_max_256(float v, float* sm /*[8]*/) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = sm[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) t = fmaxf(t, sm[i]);
  return t;
}

}  // namespace coopcap
