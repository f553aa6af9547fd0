// Additive attention over packed region features (AttModel.py:465-489), forward / per-step
// backward / deferred region-gradient accumulation.
//
// All three kernels are HBM-bound streams over the region tensors of one batch row per CTA:
// chunks of 8 regions (8 x AR bf16 = 8 KB per tensor) are brought into shared memory by TMA bulk
// copies (cp.async.bulk + mbarrier transaction counts) through a multi-stage ring, so the copy
// engine keeps several KB per CTA in flight while the warps compute on the previous chunk.
//   forward : one pass over (p_att, att_e) with an online softmax (running max / sum / weighted sum)
//   backward: pass 1 over att_e (dw_l = <d_att_res, att_e_l>), softmax backward, pass 2 over p_att
//             (d_att_h = alpha * sum_l de_l (1 - tanh^2))
//   deferred: one pass over p_att for all steps at once (d_p_att, d_att_e, alpha / bias partials)
#pragma once
#include "common.cuh"
#include "speaker_kernels.cuh"

namespace coopcap {

constexpr int ATT_THREADS = 256;   // 8 warps
constexpr int ATT_CH = 8;          // regions per chunk (one per warp in the score phase)

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int AR, int STAGES>
__global__ void __launch_bounds__(ATT_THREADS)
attention_fwd2_kernel(const __nv_bfloat16* __restrict__ p_att16, const __nv_bfloat16* __restrict__ att_e16,
                      const int* __restrict__ off, int Lfix, const float* __restrict__ s_row0,
                      int64_t lds, int att_h_col, const float* __restrict__ w_alpha,
                      __nv_bfloat16* __restrict__ att_res16, float* __restrict__ att_w) {
  static_assert(AR == 2 * ATT_THREADS, "thread -> 2 columns mapping");
  constexpr int EPL = AR / 32;                    // score elements per lane (16)
  constexpr int CHUNK_BYTES = ATT_CH * AR * 2;    // one tensor, one chunk
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stages = smem;                                             // [STAGES][2][CHUNK_BYTES]
  float* s_e = reinterpret_cast<float*>(smem + STAGES * 2 * CHUNK_BYTES);   // [Lb] scores
  const int b = blockIdx.x;
  const int r0 = off ? off[b] : b * Lfix;
  const int Lb = off ? off[b + 1] - r0 : Lfix;
  float* s_ce = s_e + ((Lb + 7) & ~7);                                // [ATT_CH]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ce + ATT_CH);        // [STAGES]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = (Lb + ATT_CH - 1) / ATT_CH;
  const __nv_bfloat16* pg = p_att16 + int64_t(r0) * AR;
  const __nv_bfloat16* eg = att_e16 + int64_t(r0) * AR;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  auto issue = [&](int c) {
    const int st = c % STAGES;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
    const uint32_t bytes = uint32_t(rows) * AR * 2;
    mbar_expect_tx(&bars[st], 2 * bytes);
    bulk_load_1d(stages + (st * 2 + 0) * CHUNK_BYTES, pg + int64_t(c) * ATT_CH * AR, bytes, &bars[st]);
    bulk_load_1d(stages + (st * 2 + 1) * CHUNK_BYTES, eg + int64_t(c) * ATT_CH * AR, bytes, &bars[st]);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < min(STAGES, nch); ++c) issue(c);

  // per-lane slices of att_h and alpha for the score phase: columns lane*8.. and AR/2 + lane*8..
  float ah[EPL], al[EPL];
  {
    const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
#pragma unroll
    for (int h = 0; h < EPL / 8; ++h)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = h * 256 + lane * 8 + j;
        ah[h * 8 + j] = att_h[col];
        al[h * 8 + j] = __ldg(w_alpha + col);
      }
  }
  float m = -INFINITY, sum = 0.f, acc0 = 0.f, acc1 = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int st = c % STAGES;
    mbar_wait(&bars[st], (c / STAGES) & 1);
    const uint8_t* ps = stages + (st * 2 + 0) * CHUNK_BYTES;
    const uint8_t* es = stages + (st * 2 + 1) * CHUNK_BYTES;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
    // scores: warp w -> region c*8 + w
    {
      float e = -INFINITY;
      if (warp < rows) {
        float a = 0.f;
#pragma unroll
        for (int h = 0; h < EPL / 8; ++h) {
          const uint4 u = *reinterpret_cast<const uint4*>(ps + warp * (AR * 2) + (h * 256 + lane * 8) * 2);
          float f[8];
          bf16x8_to_float(u, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) a += al[h * 8 + j] * tanh_fast(f[j] + ah[h * 8 + j]);
        }
        e = warp_sum(a);
      }
      if (lane == 0) {
        s_ce[warp] = e;
        if (warp < rows) s_e[c * ATT_CH + warp] = e;
      }
    }
    __syncthreads();
    // online softmax update; thread -> columns 2*tid, 2*tid+1
    {
      float ce[ATT_CH];
      float cm = m;
#pragma unroll
      for (int w = 0; w < ATT_CH; ++w) { ce[w] = s_ce[w]; cm = fmaxf(cm, ce[w]); }
      const float scale = __expf(m - cm);      // exp(-inf) = 0 on the first chunk
      m = cm;
      sum *= scale; acc0 *= scale; acc1 *= scale;
#pragma unroll
      for (int w = 0; w < ATT_CH; ++w) {
        if (w < rows) {
          const float pw = __expf(ce[w] - m);
          const float2 v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(es + w * (AR * 2) + threadIdx.x * 4));
          sum += pw;
          acc0 += pw * v.x;
          acc1 += pw * v.y;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && c + STAGES < nch) issue(c + STAGES);
  }
  const float inv = 1.f / sum;
  {
    __nv_bfloat162 o = __floats2bfloat162_rn(acc0 * inv, acc1 * inv);
    *reinterpret_cast<__nv_bfloat162*>(att_res16 + int64_t(b) * AR + 2 * threadIdx.x) = o;
  }
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) att_w[r0 + l] = __expf(s_e[l] - m) * inv;
}

template <int AR, int STAGES>
size_t attention_fwd2_smem(int L) {
  return size_t(STAGES) * 2 * ATT_CH * AR * 2 + sizeof(float) * (((L + 7) & ~7) + ATT_CH) +
         sizeof(uint64_t) * STAGES + 16;
}

// ------------------------------------------------------------------------------------------
// per-step backward (inside the BPTT chain): d(scores) and d(att_h)
// ------------------------------------------------------------------------------------------
template <int AR, int STAGES>
__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd2_kernel(const __nv_bfloat16* __restrict__ p_att16, const __nv_bfloat16* __restrict__ att_e16,
                      const int* __restrict__ off, int Lfix, const float* __restrict__ s_row0,
                      int64_t lds, int att_h_col, const float* __restrict__ w_alpha,
                      const float* __restrict__ d_att_res, const float* __restrict__ att_w,
                      float* __restrict__ de_out, __nv_bfloat16* __restrict__ dscat) {
  static_assert(AR == 2 * ATT_THREADS, "thread -> 2 columns mapping");
  constexpr int EPL = AR / 32;
  constexpr int CHUNK_BYTES = ATT_CH * AR * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stages = smem;                                              // [STAGES][CHUNK_BYTES]
  float* s_de = reinterpret_cast<float*>(smem + STAGES * CHUNK_BYTES); // [Lb] dw -> de
  const int b = blockIdx.x;
  const int r0 = off ? off[b] : b * Lfix;
  const int Lb = off ? off[b + 1] - r0 : Lfix;
  float* s_red = s_de + ((Lb + 7) & ~7);                               // [8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = (Lb + ATT_CH - 1) / ATT_CH;
  const __nv_bfloat16* pg = p_att16 + int64_t(r0) * AR;
  const __nv_bfloat16* eg = att_e16 + int64_t(r0) * AR;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  // item i in [0, 2*nch): att_e chunks first, then p_att chunks
  auto issue = [&](int i) {
    const int st = i % STAGES;
    const int c = i < nch ? i : i - nch;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
    const uint32_t bytes = uint32_t(rows) * AR * 2;
    mbar_expect_tx(&bars[st], bytes);
    bulk_load_1d(stages + st * CHUNK_BYTES, (i < nch ? eg : pg) + int64_t(c) * ATT_CH * AR, bytes,
                 &bars[st]);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < min(STAGES, 2 * nch); ++i) issue(i);
  // pass 1: dw_l = <d_att_res, att_e_l>; lane slices of d_att_res in registers
  float dr[EPL];
#pragma unroll
  for (int h = 0; h < EPL / 8; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) dr[h * 8 + j] = d_att_res[int64_t(b) * AR + h * 256 + lane * 8 + j];
  for (int c = 0; c < nch; ++c) {
    const int st = c % STAGES;
    mbar_wait(&bars[st], (c / STAGES) & 1);
    const uint8_t* es = stages + st * CHUNK_BYTES;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
    if (warp < rows) {
      float a = 0.f;
#pragma unroll
      for (int h = 0; h < EPL / 8; ++h) {
        const uint4 u = *reinterpret_cast<const uint4*>(es + warp * (AR * 2) + (h * 256 + lane * 8) * 2);
        float f[8];
        bf16x8_to_float(u, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) a += dr[h * 8 + j] * f[j];
      }
      a = warp_sum(a);
      if (lane == 0) s_de[c * ATT_CH + warp] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0 && c + STAGES < 2 * nch) issue(c + STAGES);
  }
  // softmax backward: de_l = w_l (dw_l - sum_l' w_l' dw_l')
  float dotw = 0.f;
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) dotw += att_w[r0 + l] * s_de[l];
  dotw = block_sum_256(dotw, s_red);
  __syncthreads();
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) {
    const float de = att_w[r0 + l] * (s_de[l] - dotw);
    s_de[l] = de;
    de_out[r0 + l] = de;
  }
  __syncthreads();
  // pass 2: d_att_h[j] = alpha_j sum_l de_l (1 - tanh^2(p_att[l,j] + att_h[j])); thread -> 2 columns
  const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
  const float ah0 = att_h[2 * threadIdx.x], ah1 = att_h[2 * threadIdx.x + 1];
  float acc0 = 0.f, acc1 = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int i = nch + c;
    const int st = i % STAGES;
    mbar_wait(&bars[st], (i / STAGES) & 1);
    const uint8_t* ps = stages + st * CHUNK_BYTES;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
#pragma unroll
    for (int w = 0; w < ATT_CH; ++w) {
      if (w < rows) {
        const float de = s_de[c * ATT_CH + w];
        const float2 v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(ps + w * (AR * 2) + threadIdx.x * 4));
        const float t0 = tanh_fast(v.x + ah0), t1 = tanh_fast(v.y + ah1);
        acc0 += de * (1.f - t0 * t0);
        acc1 += de * (1.f - t1 * t1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && i + STAGES < 2 * nch) issue(i + STAGES);
  }
  {
    const float a0 = __ldg(w_alpha + 2 * threadIdx.x), a1 = __ldg(w_alpha + 2 * threadIdx.x + 1);
    __nv_bfloat162 o = __floats2bfloat162_rn(acc0 * a0, acc1 * a1);
    *reinterpret_cast<__nv_bfloat162*>(dscat + int64_t(b) * lds + att_h_col + 2 * threadIdx.x) = o;
  }
}

template <int AR, int STAGES>
size_t attention_bwd2_smem(int L) {
  return size_t(STAGES) * ATT_CH * AR * 2 + sizeof(float) * (((L + 7) & ~7) + 8) +
         sizeof(uint64_t) * STAGES + 16;
}

// ------------------------------------------------------------------------------------------
// deferred accumulation over all steps (after the BPTT loop), one pass over p_att:
//   d_att_e[l,:]  = sum_t w_t[l] * d_att_res_t[:]
//   d_p_att[l,j]  = alpha_j * sum_t de_t[l] (1 - tanh^2(p_att[l,j] + att_h_t[j]))
//   galpha[b, j]  = sum_t sum_l de_t[l] tanh(p_att[l,j] + att_h_t[j])   (per-row partial)
//   gbias[b, j]   = sum_l d_p_att[l,j]                                  (per-row partial, fp32)
// thread -> 2 columns x the 8 regions of a chunk; the per-step vectors are read once per chunk.
// ------------------------------------------------------------------------------------------
template <int AR, int STAGES>
__global__ void __launch_bounds__(ATT_THREADS)
attention_deferred2_kernel(const __nv_bfloat16* __restrict__ p_att16, const int* __restrict__ off,
                           int Lfix, const float* __restrict__ s_all, int64_t lds, int att_h_col,
                           int64_t step_stride_s, const float* __restrict__ w_alpha,
                           const float* __restrict__ d_att_res, const float* __restrict__ att_w,
                           const float* __restrict__ de, int NL, int n_steps, int B,
                           float* __restrict__ d_att_e, __nv_bfloat16* __restrict__ d_p_att16,
                           float* __restrict__ galpha_part, float* __restrict__ gbias_part) {
  static_assert(AR == 2 * ATT_THREADS, "thread -> 2 columns mapping");
  constexpr int CHUNK_BYTES = ATT_CH * AR * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int b = blockIdx.x;
  const int r0 = off ? off[b] : b * Lfix;
  const int Lb = off ? off[b + 1] - r0 : Lfix;
  const int Lp = (Lb + 7) & ~7;
  uint8_t* stages = smem;                                                   // [STAGES][CHUNK_BYTES]
  float* s_ah = reinterpret_cast<float*>(smem + STAGES * CHUNK_BYTES);      // [n][AR]
  float* s_dr = s_ah + n_steps * AR;                                        // [n][AR]
  float* s_w = s_dr + n_steps * AR;                                         // [n][Lp]
  float* s_de = s_w + n_steps * Lp;                                         // [n][Lp]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_de + n_steps * Lp);
  const int nch = (Lb + ATT_CH - 1) / ATT_CH;
  const __nv_bfloat16* pg = p_att16 + int64_t(r0) * AR;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  auto issue = [&](int c) {
    const int st = c % STAGES;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
    const uint32_t bytes = uint32_t(rows) * AR * 2;
    mbar_expect_tx(&bars[st], bytes);
    bulk_load_1d(stages + st * CHUNK_BYTES, pg + int64_t(c) * ATT_CH * AR, bytes, &bars[st]);
  };
  if (threadIdx.x == 0)
    for (int c = 0; c < min(STAGES, nch); ++c) issue(c);
  for (int i = threadIdx.x; i < n_steps * AR; i += ATT_THREADS) {
    const int t = i / AR, j = i % AR;
    s_ah[i] = s_all[int64_t(t) * step_stride_s + int64_t(b) * lds + att_h_col + j];
    s_dr[i] = d_att_res[(int64_t(t) * B + b) * AR + j];
  }
  for (int i = threadIdx.x; i < n_steps * Lp; i += ATT_THREADS) {
    const int t = i / Lp, l = i % Lp;
    s_w[i] = l < Lb ? att_w[int64_t(t) * NL + r0 + l] : 0.f;
    s_de[i] = l < Lb ? de[int64_t(t) * NL + r0 + l] : 0.f;
  }
  __syncthreads();
  const int j0 = 2 * threadIdx.x;
  const float al0 = __ldg(w_alpha + j0), al1 = __ldg(w_alpha + j0 + 1);
  float ga0 = 0.f, ga1 = 0.f, gb0 = 0.f, gb1 = 0.f;
  for (int c = 0; c < nch; ++c) {
    const int st = c % STAGES;
    mbar_wait(&bars[st], (c / STAGES) & 1);
    const uint8_t* ps = stages + st * CHUNK_BYTES;
    const int rows = min(ATT_CH, Lb - c * ATT_CH);
    float p0[ATT_CH], p1[ATT_CH], ap0[ATT_CH], ap1[ATT_CH], ae0[ATT_CH], ae1[ATT_CH];
#pragma unroll
    for (int w = 0; w < ATT_CH; ++w) {
      float2 v = make_float2(0.f, 0.f);
      if (w < rows) v = bf2_to_f2(*reinterpret_cast<const uint32_t*>(ps + w * (AR * 2) + threadIdx.x * 4));
      p0[w] = v.x; p1[w] = v.y;
      ap0[w] = ap1[w] = ae0[w] = ae1[w] = 0.f;
    }
    for (int t = 0; t < n_steps; ++t) {
      const float2 ah = *reinterpret_cast<const float2*>(s_ah + t * AR + j0);
      const float2 dr = *reinterpret_cast<const float2*>(s_dr + t * AR + j0);
      const float4 d_lo = *reinterpret_cast<const float4*>(s_de + t * Lp + c * ATT_CH);
      const float4 d_hi = *reinterpret_cast<const float4*>(s_de + t * Lp + c * ATT_CH + 4);
      const float4 w_lo = *reinterpret_cast<const float4*>(s_w + t * Lp + c * ATT_CH);
      const float4 w_hi = *reinterpret_cast<const float4*>(s_w + t * Lp + c * ATT_CH + 4);
      const float d8[8] = {d_lo.x, d_lo.y, d_lo.z, d_lo.w, d_hi.x, d_hi.y, d_hi.z, d_hi.w};
      const float w8[8] = {w_lo.x, w_lo.y, w_lo.z, w_lo.w, w_hi.x, w_hi.y, w_hi.z, w_hi.w};
#pragma unroll
      for (int w = 0; w < ATT_CH; ++w) {
        const float t0 = tanh_fast(p0[w] + ah.x), t1 = tanh_fast(p1[w] + ah.y);
        ap0[w] += d8[w] * (1.f - t0 * t0);
        ap1[w] += d8[w] * (1.f - t1 * t1);
        ga0 += d8[w] * t0;
        ga1 += d8[w] * t1;
        ae0[w] += w8[w] * dr.x;
        ae1[w] += w8[w] * dr.y;
      }
    }
#pragma unroll
    for (int w = 0; w < ATT_CH; ++w) {
      if (w < rows) {
        const int64_t row = int64_t(r0 + c * ATT_CH + w);
        const float o0 = ap0[w] * al0, o1 = ap1[w] * al1;
        gb0 += o0; gb1 += o1;
        *reinterpret_cast<__nv_bfloat162*>(d_p_att16 + row * AR + j0) = __floats2bfloat162_rn(o0, o1);
        *reinterpret_cast<float2*>(d_att_e + row * AR + j0) = make_float2(ae0[w], ae1[w]);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && c + STAGES < nch) issue(c + STAGES);
  }
  *reinterpret_cast<float2*>(galpha_part + int64_t(b) * AR + j0) = make_float2(ga0, ga1);
  *reinterpret_cast<float2*>(gbias_part + int64_t(b) * AR + j0) = make_float2(gb0, gb1);
}

template <int AR, int STAGES>
size_t attention_deferred2_smem(int L, int n_steps) {
  const int Lp = (L + 7) & ~7;
  return size_t(STAGES) * ATT_CH * AR * 2 + sizeof(float) * (2 * size_t(n_steps) * AR + 2 * size_t(n_steps) * Lp) +
         sizeof(uint64_t) * STAGES + 16;
}

}  // namespace coopcap
