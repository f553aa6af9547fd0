// Speaker (Att2in2) backward: logit gradients of the straight-through samplers and of log-prob
// losses, BPTT through the decode loop, deferred accumulation of the region-tensor gradients and
// all weight gradients.  Maths: SURVEY.md Appendix A.2/A.3/A.5; buffers: include/coopcap.h.
#include <algorithm>
#include "../../include/coopcap.h"
#include "common.cuh"
#include "gemm.cuh"
#include "speaker_kernels.cuh"
#include "attention.cuh"

namespace coopcap {

using bf16 = __nv_bfloat16;

size_t attention_smem_bytes(int A, int R, int L);

// ------------------------------------------------------------------------------------------
// small shared helpers
// ------------------------------------------------------------------------------------------
// out[c] = sum_r src[r, c]   (bf16 or fp32 source); out is overwritten.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ src, int64_t rows, int cols, int64_t ld,
                              float* __restrict__ out, int64_t rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (c < cols) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      if constexpr (sizeof(T) == 2) acc += __bfloat162float(src[r * ld + c]);
      else acc += src[r * ld + c];
    }
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// bf16 source with 16-byte aligned rows: every thread owns 8 adjacent columns (one 16-byte load per
// row, a warp reads 512 contiguous bytes), four rows in flight per thread
__global__ void __launch_bounds__(256)
colsum_bf16x8_kernel(const bf16* __restrict__ src, int64_t rows, int cols, int64_t ld,
                     float* __restrict__ out, int64_t rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[8][32][9];
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < cols) {
    int64_t r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const uint4*>(src + (r + 8 * q) * ld + c0));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
        bf16x8_to_float(v[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    for (; r < r1; r += 8) {
      float f[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(src + r * ld + c0)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  // 256 threads reduce the 8 row-slices of the block's 256 columns
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int col = blockIdx.x * 256 + t;
  if (col < cols) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += sm[i][t >> 3][t & 7];
    atomicAdd(out + col, tot);
  }
}

template <typename T>
static int colsum_impl(const T* src, int64_t rows, int cols, int64_t ld, float* out, cudaStream_t s) {
  CC_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
  if (rows <= 0) return CC_OK;
  if constexpr (sizeof(T) == 2) {
    if (cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int col_blocks = (cols + 255) / 256;
      int64_t row_blocks = (int64_t(num_sms()) * 4 + col_blocks - 1) / col_blocks;
      if (row_blocks > (rows + 63) / 64) row_blocks = (rows + 63) / 64;
      if (row_blocks < 1) row_blocks = 1;
      const int64_t rpb = (rows + row_blocks - 1) / row_blocks;
      CC_CHECK_CUDA(launch_pdl(colsum_bf16x8_kernel, dim3(col_blocks, (unsigned)row_blocks), dim3(32, 8), size_t(0), s, 
          reinterpret_cast<const bf16*>(src), rows, cols, ld, out, rpb));
      CC_LAUNCH_CHECK_K(PROF_REDUCE, s, 0.0, 0.0);
      return CC_OK;
    }
  }
  const int col_blocks = (cols + 31) / 32;
  int64_t row_blocks = (int64_t(num_sms()) * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > (rows + 63) / 64) row_blocks = (rows + 63) / 64;
  if (row_blocks < 1) row_blocks = 1;
  const int64_t rpb = (rows + row_blocks - 1) / row_blocks;
  CC_CHECK_CUDA(launch_pdl(colsum_kernel<T>, dim3(col_blocks, (unsigned)row_blocks), dim3(32, 8), size_t(0), s, src, rows, cols,
                                                                                  ld, out, rpb));
  CC_LAUNCH_CHECK_K(PROF_REDUCE, s, 0.0, 0.0);
  return CC_OK;
}
int colsum_bf16(const void* src, int64_t rows, int cols, int64_t ld, float* out, cudaStream_t s) {
  return colsum_impl(reinterpret_cast<const bf16*>(src), rows, cols, ld, out, s);
}
int colsum_f32(const float* src, int64_t rows, int cols, int64_t ld, float* out, cudaStream_t s) {
  return colsum_impl(src, rows, cols, ld, out, s);
}

// weight-gradient contraction C[M,N] = A^T B with A stored [K, M], B stored [K, N] (bf16),
// split along K so that the grid fills the machine; C is overwritten.
int wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                 float* C, int64_t ldc, cudaStream_t s) {
  const int bn = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);  // split-K sizing at the widest tile
  int split = num_sms() / tiles;
  const int max_split = (K + 255) / 256;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  EpiStoreParams e = {};
  e.alpha = 1.f; e.C = C; e.ldc = ldc;
  if (split > 1) {
    e.mode = 2;
    CC_CHECK_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, N * sizeof(float), M, s));
  }
  return gemm_run(0, 1, 1, A, lda, B, ldb, M, N, K, split, bn, e, s);
}

// ------------------------------------------------------------------------------------------
// d(loss)/d(logits) of the straight-through samplers, one CTA per row:
//   y = softmax(score), score = (z+G)/tau (gumbel) or z/tau (multinomial), rebuilt from the saved
//   (max, sum) and the regenerated / injected noise;  dz = y (g - <y,g>) / tau on unfinished rows.
// ------------------------------------------------------------------------------------------
// GT = float (dense upstream gradient handed in by a foreign caller, partial-sampling scratch) or
// bf16 (the factored path's demb . W_emb^T, which never needs more: dz is stored in bf16 anyway)
template <typename GT>
__device__ __forceinline__ void ld_g4(const GT* p, float (&o)[4]);
template <>
__device__ __forceinline__ void ld_g4<float>(const float* p, float (&o)[4]) {
  o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3];
}
template <>
__device__ __forceinline__ void ld_g4<bf16>(const bf16* p, float (&o)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}

// PERT: z holds the perturbed logits z + G (coopcap_speaker.store_perturbed): no noise is regenerated
template <typename GT, bool PERT = false>
__global__ void __launch_bounds__(256)
st_bwd_kernel(const __half* __restrict__ z, const GT* __restrict__ g, int64_t ldg, int V1, int mode,
              float inv_tau, const float* __restrict__ noise, uint64_t seed, uint64_t nstream0,
              int B, const float* __restrict__ ymax, const float* __restrict__ ysum,
              const uint8_t* __restrict__ unf, bf16* __restrict__ dz,
              const float* __restrict__ lse) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float s_y[];   // [V1]
  __shared__ float red[8];
  // rows are (step, batch row) pairs of consecutive steps: row = (t - t0) * B + b
  const int64_t row = blockIdx.x;
  const int b = int(row % B);
  const uint64_t nstream = nstream0 + uint64_t(row / B);
  bf16* dr = dz + row * V1;
  if (!unf[row]) {
    for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256) store_bf16x4(dr + 4 * v4, 0.f, 0.f, 0.f, 0.f);
    return;
  }
  const __half* zr = z + row * V1;
  const GT* gr = g + row * ldg;
  const float* nr = noise ? noise + row * V1 : nullptr;
  const float m = ymax[row], inv_s = 1.f / ysum[row];
  const bool fast = (noise == nullptr);
  float dot = 0.f;
  for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256) {
    float x4[4];
    f16x4_to_float(zr + 4 * v4, x4);
    float g4[4];
    ld_g4<GT>(gr + 4 * v4, g4);
    float u4[4] = {0.f, 0.f, 0.f, 0.f};
    if (!PERT && (mode == COOPCAP_SAMPLE_ST_GUMBEL || mode == COOPCAP_SAMPLE_PS_GUMBEL))
      noise4(nr, v4, seed, nstream, uint64_t(b) * (V1 / 4) + v4, u4);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float score = PERT ? x4[q] * inv_tau : st_score(mode, x4[q], u4[q], inv_tau, fast);
      const float y = ex2_ftz((score - m) * 1.4426950408889634f) * inv_s;
      s_y[4 * v4 + q] = y;
      dot += y * g4[q];
    }
  }
  dot = block_sum_256(dot, red);
  if (mode == COOPCAP_SAMPLE_PS_MULTINOMIAL) {
    // y = exp(lp / tau) with lp = log_softmax(z):  d lp = y g / tau,  dz = d lp - softmax(z) sum(d lp)
    const float l = lse[row];
    for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256) {
      const float4 yv = *reinterpret_cast<const float4*>(s_y + 4 * v4);
      float zq[4];
      f16x4_to_float(zr + 4 * v4, zq);
      const float p0 = ex2_ftz((zq[0] - l) * 1.4426950408889634f), p1 = ex2_ftz((zq[1] - l) * 1.4426950408889634f);
      const float p2 = ex2_ftz((zq[2] - l) * 1.4426950408889634f), p3 = ex2_ftz((zq[3] - l) * 1.4426950408889634f);
      float g4[4];
      ld_g4<GT>(gr + 4 * v4, g4);
      store_bf16x4(dr + 4 * v4, inv_tau * (yv.x * g4[0] - p0 * dot), inv_tau * (yv.y * g4[1] - p1 * dot),
                   inv_tau * (yv.z * g4[2] - p2 * dot), inv_tau * (yv.w * g4[3] - p3 * dot));
    }
    return;
  }
  for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256) {
    const float4 yv = *reinterpret_cast<const float4*>(s_y + 4 * v4);
    float g4[4];
    ld_g4<GT>(gr + 4 * v4, g4);
    store_bf16x4(dr + 4 * v4, inv_tau * yv.x * (g4[0] - dot), inv_tau * yv.y * (g4[1] - dot),
                 inv_tau * yv.z * (g4[2] - dot), inv_tau * yv.w * (g4[3] - dot));
  }
}

// The benchmarked case of st_bwd_kernel -- perturbed fp16 logits z + G kept by the forward pass, bf16
// upstream gradient from the factored listener path -- with the row held in REGISTERS: every thread
// owns up to ST_REG_IT groups of 8 consecutive columns, issues all its 16-byte loads of z and g
// up front, forms y and <y, g>, and writes dz from the same registers.  One read of each input, one
// write, no shared-memory copy of y, no second pass over g.  V1 <= 8 * 256 * ST_REG_IT.
constexpr int ST_REG_IT = 5;
__global__ void __launch_bounds__(256)
st_bwd_pert_kernel(const __half* __restrict__ z, const bf16* __restrict__ g, int V1, float inv_tau,
                   const float* __restrict__ ymax, const float* __restrict__ ysum,
                   const uint8_t* __restrict__ unf, bf16* __restrict__ dz) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const int64_t row = blockIdx.x;
  const int nv8 = V1 >> 3;
  uint4* dr = reinterpret_cast<uint4*>(dz + row * V1);
  if (!unf[row]) {
    for (int v8 = threadIdx.x; v8 < nv8; v8 += 256) dr[v8] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const uint4* zr = reinterpret_cast<const uint4*>(z + row * V1);
  const uint4* gr = reinterpret_cast<const uint4*>(g + row * V1);
  uint4 zq[ST_REG_IT], gq[ST_REG_IT];
#pragma unroll
  for (int k = 0; k < ST_REG_IT; ++k) {
    const int v8 = threadIdx.x + k * 256;
    if (v8 < nv8) { zq[k] = zr[v8]; gq[k] = gr[v8]; }
  }
  const float m = ymax[row], inv_s = 1.f / ysum[row];
  float y[ST_REG_IT][8];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < ST_REG_IT; ++k) {
    const int v8 = threadIdx.x + k * 256;
    if (v8 < nv8) {
      float x8[8], g8[8];
      f16x8_to_float(zq[k], x8);
      bf16x8_to_float(gq[k], g8);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        y[k][q] = ex2_ftz((x8[q] * inv_tau - m) * 1.4426950408889634f) * inv_s;
        dot += y[k][q] * g8[q];
      }
    }
  }
  dot = block_sum_256(dot, red);
#pragma unroll
  for (int k = 0; k < ST_REG_IT; ++k) {
    const int v8 = threadIdx.x + k * 256;
    if (v8 < nv8) {
      float g8[8], o[8];
      bf16x8_to_float(gq[k], g8);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = inv_tau * y[k][q] * (g8[q] - dot);
      dr[v8] = float8_to_bf16x8(o);
    }
  }
}

// d(loss)/d(v . embed) of the partial-sampling next input x = dropout(relu(v . embed)): the ReLU
// and dropout decisions are both recovered from the saved x (x > 0)
__global__ void ps_dpre_kernel(const float* __restrict__ d_xh, const bf16* __restrict__ xh16, int E,
                               int XH, float scale, bf16* __restrict__ dpre16) {
  const int64_t row = blockIdx.x;
  for (int i = threadIdx.x; i < E; i += blockDim.x)
    dpre16[row * E + i] = __float2bfloat16_rn(
        __bfloat162float(xh16[row * XH + i]) > 0.f ? d_xh[row * XH + i] * scale : 0.f);
}

// dz = coef * (onehot(tok) - softmax(z))   (REINFORCE / XE), one CTA per (step, row); ACC: added to
// the gradient already in dz (a second loss term on the same pass, e.g. the CIDEr term next to the
// straight-through listener gradient, AlternatingJointModel.py:490-503)
template <bool ACC>
__global__ void __launch_bounds__(256)
logp_bwd_kernel(const __half* __restrict__ z, int V1, const float* __restrict__ lse,
                const int64_t* __restrict__ tok, const float* __restrict__ coef,
                bf16* __restrict__ dz) {
  const int64_t row = blockIdx.x;
  bf16* dr = dz + row * V1;
  const float c = coef[row];
  if (c == 0.f) {
    if (!ACC)
      for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256) store_bf16x4(dr + 4 * v4, 0.f, 0.f, 0.f, 0.f);
    return;
  }
  const __half* zr = z + row * V1;
  const float l = lse[row];
  const int t = int(tok[row]);
  for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256) {
    float x4[4];
    f16x4_to_float(zr + 4 * v4, x4);
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = c * ((4 * v4 + q == t ? 1.f : 0.f) - ex2_ftz((x4[q] - l) * 1.4426950408889634f));
    if (ACC) {
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(dr + 4 * v4);
      const float2 a = __bfloat1622float2(p[0]), b = __bfloat1622float2(p[1]);
      o[0] += a.x; o[1] += a.y; o[2] += b.x; o[3] += b.y;
    }
    store_bf16x4(dr + 4 * v4, o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// maxout-LSTM pointwise backward (one step)
// ------------------------------------------------------------------------------------------
__global__ void lstm_bwd_kernel(const float* __restrict__ d_out, const bf16* __restrict__ out16,
                                const float* __restrict__ dh_next, int64_t ld_dh,
                                const float* __restrict__ dc_in, float* __restrict__ dc_out,
                                const float* __restrict__ s, int64_t lds, const float* __restrict__ u,
                                const float* __restrict__ c_prev, const float* __restrict__ c_cur,
                                bf16* __restrict__ dscat, float drop_p, int B, int R,
                                float* __restrict__ zero_row0, int zero_cols) {
  pdl_launch_dependents();
  pdl_wait();
  // 4 hidden units per thread: 16-byte loads, 8-byte bf16 stores
  const int idx4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = R / 4;
  if (idx4 >= B * per_row) return;
  const int b = idx4 / per_row, j = (idx4 % per_row) * 4;
  if (zero_row0) {
    // this step's slice of d[x|h]: the split-K dgrad GEMM that follows reduce-ADDs into it (was one
    // 67 MB memset in front of the loop); zero_cols floats per row, spread over the row's threads
    float4* zr = reinterpret_cast<float4*>(zero_row0 + int64_t(b) * zero_cols);
    for (int q = idx4 % per_row; q < zero_cols / 4; q += per_row) zr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int64_t idx = int64_t(b) * R + j;
  auto ld4 = [](const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  };
  float dh[4], si[4], sf[4], so[4], s1[4], s2[4], u1[4], u2[4], cc[4], cp[4], dci[4] = {0.f, 0.f, 0.f, 0.f};
  ld4(d_out + idx, dh);
  const float* sr = s + int64_t(b) * lds;
  ld4(sr + j, si); ld4(sr + R + j, sf); ld4(sr + 2 * R + j, so); ld4(sr + 3 * R + j, s1);
  ld4(sr + 4 * R + j, s2);
  ld4(u + int64_t(b) * 2 * R + j, u1); ld4(u + int64_t(b) * 2 * R + R + j, u2);
  ld4(c_cur + idx, cc); ld4(c_prev + idx, cp);
  if (dc_in) ld4(dc_in + idx, dci);
  if (drop_p > 0.f) {
    const uint2 o = *reinterpret_cast<const uint2*>(out16 + idx);
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&o.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&o.y);
    const float k[4] = {__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi)};
    const float sc = 1.f / (1.f - drop_p);
#pragma unroll
    for (int q = 0; q < 4; ++q) dh[q] = (k[q] != 0.f) ? dh[q] * sc : 0.f;
  }
  if (dh_next) {
    float n[4];
    ld4(dh_next + int64_t(b) * ld_dh + j, n);
#pragma unroll
    for (int q = 0; q < 4; ++q) dh[q] += n[q];
  }
  float o_dc[4], o_i[4], o_f[4], o_o[4], o_g1[4], o_g2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float ig = 1.f / (1.f + expf(-si[q]));
    const float fg = 1.f / (1.f + expf(-sf[q]));
    const float og = 1.f / (1.f + expf(-so[q]));
    const float a1 = s1[q] + u1[q];
    const float a2 = s2[q] + u2[q];
    const float g = fmaxf(a1, a2);
    const float tc = tanhf(cc[q]);
    const float dc = dci[q] + dh[q] * og * (1.f - tc * tc);
    const float d_o = dh[q] * tc;
    const float d_f = dc * cp[q];
    const float d_i = dc * g;
    const float d_g = dc * ig;
    o_dc[q] = dc * fg;
    o_i[q] = d_i * ig * (1.f - ig);
    o_f[q] = d_f * fg * (1.f - fg);
    o_o[q] = d_o * og * (1.f - og);
    const bool first = a1 >= a2;
    o_g1[q] = first ? d_g : 0.f;
    o_g2[q] = first ? 0.f : d_g;
  }
  *reinterpret_cast<float4*>(dc_out + idx) = make_float4(o_dc[0], o_dc[1], o_dc[2], o_dc[3]);
  bf16* dr = dscat + int64_t(b) * lds;
  store_bf16x4(dr + j, o_i[0], o_i[1], o_i[2], o_i[3]);
  store_bf16x4(dr + R + j, o_f[0], o_f[1], o_f[2], o_f[3]);
  store_bf16x4(dr + 2 * R + j, o_o[0], o_o[1], o_o[2], o_o[3]);
  store_bf16x4(dr + 3 * R + j, o_g1[0], o_g1[1], o_g1[2], o_g1[3]);
  store_bf16x4(dr + 4 * R + j, o_g2[0], o_g2[1], o_g2[2], o_g2[3]);
}

// ------------------------------------------------------------------------------------------
// additive attention backward for one step (inside the BPTT chain): produces d(scores) and
// d(att_h); the region-tensor gradients are accumulated over all steps by the deferred kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd_kernel(const bf16* __restrict__ p_att16, const bf16* __restrict__ att_e16,
                     const int* __restrict__ off, int Lfix, const float* __restrict__ s_row0,
                     int64_t lds, int att_h_col, const float* __restrict__ w_alpha,
                     const float* __restrict__ d_att_res, const float* __restrict__ att_w,
                     float* __restrict__ de_out, bf16* __restrict__ dscat, int A, int R) {
  extern __shared__ float sm[];
  float* s_ah = sm;              // [A]
  float* s_dr = sm + A;          // [R] d_att_res   (A == R is not assumed: sized 2*max below)
  const int mx = A > R ? A : R;
  s_dr = sm + mx;
  float* s_red = sm + 2 * mx;    // [8]
  float* s_e = sm + 2 * mx + 8;  // [Lb] dw -> de
  const int b = blockIdx.x;
  const int r0 = off ? off[b] : b * Lfix;
  const int Lb = off ? off[b + 1] - r0 : Lfix;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
  for (int i = threadIdx.x; i < A; i += ATT_THREADS) s_ah[i] = att_h[i];
  for (int i = threadIdx.x; i < R; i += ATT_THREADS) s_dr[i] = d_att_res[int64_t(b) * R + i];
  __syncthreads();
  // dw_l = <d_att_res, att_e[l]>
  for (int l = warp; l < Lb; l += ATT_THREADS / 32) {
    const uint4* erow = reinterpret_cast<const uint4*>(att_e16 + int64_t(r0 + l) * R);
    float acc = 0.f;
    for (int c = lane; c < R / 8; c += 32) {
      float f[8];
      bf16x8_to_float(__ldg(erow + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += s_dr[c * 8 + j] * f[j];
    }
    acc = warp_sum(acc);
    if (lane == 0) s_e[l] = acc;
  }
  __syncthreads();
  float dotw = 0.f;
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) dotw += att_w[r0 + l] * s_e[l];
  dotw = block_sum_256(dotw, s_red);
  __syncthreads();
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) {
    const float de = att_w[r0 + l] * (s_e[l] - dotw);
    s_e[l] = de;
    de_out[r0 + l] = de;
  }
  __syncthreads();
  // d_att_h[j] = alpha_j * sum_l de_l (1 - tanh^2(p_att[l,j] + att_h[j]))
  const int tpr = A / 8, groups = ATT_THREADS / tpr;
  const int g = threadIdx.x / tpr, c = threadIdx.x % tpr;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < groups) {
    for (int l = g; l < Lb; l += groups) {
      float f[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p_att16 + int64_t(r0 + l) * A) + c), f);
      const float de = s_e[l];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float th = tanh_fast(f[j] + s_ah[c * 8 + j]);
        acc[j] += de * (1.f - th * th);
      }
    }
  }
  float* s_acc = s_e + ((Lb + 3) & ~3);  // [groups][A]
  if (g < groups) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[g * A + c * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < A; j += ATT_THREADS) {
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += s_acc[gg * A + j];
    dscat[int64_t(b) * lds + att_h_col + j] = __float2bfloat16_rn(t * w_alpha[j]);
  }
}

// Deferred accumulation over all steps, one CTA per batch row (each region tensor is read once):
//   d_att_e[l,:]  = sum_t w_t[l] * d_att_res_t[:]
//   d_p_att[l,j]  = alpha_j * sum_t de_t[l] (1 - tanh^2(p_att[l,j] + att_h_t[j]))
//   galpha[b, j]  = sum_t sum_l de_t[l] tanh(p_att[l,j] + att_h_t[j])     (per-row partial)
__global__ void __launch_bounds__(ATT_THREADS)
attention_deferred_bwd_kernel(const bf16* __restrict__ p_att16, const int* __restrict__ off, int Lfix,
                              const float* __restrict__ s_all, int64_t lds, int att_h_col,
                              int64_t step_stride_s, const float* __restrict__ w_alpha,
                              const float* __restrict__ d_att_res, const float* __restrict__ att_w,
                              const float* __restrict__ de, int NL, int n_steps, int B,
                              float* __restrict__ d_att_e, bf16* __restrict__ d_p_att16,
                              float* __restrict__ galpha_part, float* __restrict__ gbias_part,
                              int A, int R) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int r0 = off ? off[b] : b * Lfix;
  const int Lb = off ? off[b + 1] - r0 : Lfix;
  float* s_ah = sm;                              // [n_steps][A]
  float* s_dr = s_ah + n_steps * A;              // [n_steps][R]
  float* s_w = s_dr + n_steps * R;               // [n_steps][Lb]
  float* s_de = s_w + n_steps * ((Lb + 3) & ~3); // [n_steps][Lb]
  float* s_acc = s_de + n_steps * ((Lb + 3) & ~3);  // [groups][A]
  const int Lp = (Lb + 3) & ~3;
  for (int i = threadIdx.x; i < n_steps * A; i += ATT_THREADS) {
    const int t = i / A, j = i % A;
    s_ah[i] = s_all[int64_t(t) * step_stride_s + int64_t(b) * lds + att_h_col + j];
  }
  for (int i = threadIdx.x; i < n_steps * R; i += ATT_THREADS) {
    const int t = i / R, j = i % R;
    s_dr[i] = d_att_res[(int64_t(t) * B + b) * R + j];
  }
  for (int i = threadIdx.x; i < n_steps * Lb; i += ATT_THREADS) {
    const int t = i / Lb, l = i % Lb;
    s_w[t * Lp + l] = att_w[int64_t(t) * NL + r0 + l];
    s_de[t * Lp + l] = de[int64_t(t) * NL + r0 + l];
  }
  __syncthreads();
  const int tpr = A / 8, groups = ATT_THREADS / tpr;   // requires A == R (checked on the host)
  const int g = threadIdx.x / tpr, c = threadIdx.x % tpr;
  float al[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) al[j] = w_alpha[c * 8 + j];
  float ga[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < groups) {
    for (int l = g; l < Lb; l += groups) {
      float p[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(p_att16 + int64_t(r0 + l) * A) + c), p);
      float ap[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float ae[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int t = 0; t < n_steps; ++t) {
        const float w = s_w[t * Lp + l], d = s_de[t * Lp + l];
        const float* ah = s_ah + t * A + c * 8;
        const float* dr = s_dr + t * R + c * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float th = tanh_fast(p[j] + ah[j]);
          ap[j] += d * (1.f - th * th);
          ga[j] += d * th;
          ae[j] += w * dr[j];
        }
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = ap[j] * al[j]; gb[j] += o[j]; }
      *(reinterpret_cast<uint4*>(d_p_att16 + int64_t(r0 + l) * A) + c) = float8_to_bf16x8(o);
      float4* dst = reinterpret_cast<float4*>(d_att_e + int64_t(r0 + l) * R + c * 8);
      dst[0] = make_float4(ae[0], ae[1], ae[2], ae[3]);
      dst[1] = make_float4(ae[4], ae[5], ae[6], ae[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[g * A + c * 8 + j] = ga[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < A; j += ATT_THREADS) {
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += s_acc[gg * A + j];
    galpha_part[int64_t(b) * A + j] = t;
  }
  // second reduction: fp32 column sums of d_p_att (the ctx2att bias gradient cancels heavily, so
  // it is not taken from the bf16-rounded tensor)
  __syncthreads();
  if (g < groups) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[g * A + c * 8 + j] = gb[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < A; j += ATT_THREADS) {
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += s_acc[gg * A + j];
    gbias_part[int64_t(b) * A + j] = t;
  }
}

// d_pre16 = bf16(d_att_e * scale * [att_e > 0])   (relu and dropout share the "output > 0" mask)
__global__ void mask_pre_kernel(const float* __restrict__ d_att_e, const bf16* __restrict__ att_e16,
                                int64_t n4, float scale, bf16* __restrict__ d_pre16) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float4 d = *reinterpret_cast<const float4*>(d_att_e + 4 * i);
    const uint2 e = *reinterpret_cast<const uint2*>(att_e16 + 4 * i);
    const __nv_bfloat162* eh = reinterpret_cast<const __nv_bfloat162*>(&e);
    const float2 e0 = __bfloat1622float2(eh[0]), e1 = __bfloat1622float2(eh[1]);
    store_bf16x4(d_pre16 + 4 * i, e0.x > 0.f ? d.x * scale : 0.f, e0.y > 0.f ? d.y * scale : 0.f,
                 e1.x > 0.f ? d.z * scale : 0.f, e1.y > 0.f ? d.w * scale : 0.f);
  }
}

// g_embed[tok_fed[t,b], :] += d_x[t,b,:] * scale * [x16 > 0]
__global__ void embed_grad_kernel(const int64_t* __restrict__ tok_fed, const float* __restrict__ d_xh,
                                  const bf16* __restrict__ xh16, int E, int XH, float scale,
                                  float* __restrict__ g_embed) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  float* dst = g_embed + tok_fed[row] * E;
  const float* dx = d_xh + row * XH;
  const bf16* x = xh16 + row * XH;
  if ((E & 3) == 0 && (XH & 3) == 0) {
    // four columns per 16-byte vector atomic (red.global.add.v4.f32): a quarter of the L2 atomics
    for (int i = threadIdx.x * 4; i < E; i += blockDim.x * 4) {
      const float4 d = *reinterpret_cast<const float4*>(dx + i);
      const uint2 xr = *reinterpret_cast<const uint2*>(x + i);
      const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xr);
      const float2 x0 = __bfloat1622float2(xh[0]), x1 = __bfloat1622float2(xh[1]);
      const float4 v = make_float4(x0.x > 0.f ? d.x * scale : 0.f, x0.y > 0.f ? d.y * scale : 0.f,
                                   x1.x > 0.f ? d.z * scale : 0.f, x1.y > 0.f ? d.w * scale : 0.f);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
        atomicAdd(reinterpret_cast<float4*>(dst + i), v);
    }
    return;
  }
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    if (__bfloat162float(x[i]) > 0.f) atomicAdd(dst + i, dx[i] * scale);
  }
}

// ------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------
static int check_dims(const coopcap_speaker* c) {
  CC_REQUIRE(c != nullptr, "speaker: null context");
  CC_REQUIRE(c->A == c->R, "speaker backward assumes att_hid_size == rnn_size (A=%d R=%d)", c->A,
             c->R);
  CC_REQUIRE(c->n_steps >= 1 && c->n_steps <= c->cap, "speaker backward: n_steps %d", c->n_steps);
  return CC_OK;
}

int st_backward(const coopcap_speaker* c, const void* demb16, const void* w_emb16, void* g_ws,
                int g_chunk_steps, int64_t ldg, void* dz16, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  CC_REQUIRE(c->mode == COOPCAP_SAMPLE_ST_GUMBEL || c->mode == COOPCAP_SAMPLE_ST_MULTINOMIAL ||
                 c->mode == COOPCAP_SAMPLE_PS_GUMBEL || c->mode == COOPCAP_SAMPLE_PS_MULTINOMIAL,
             "st_backward: context was not sampled in a straight-through / partial-sampling mode (%d)",
             c->mode);
  const int B = c->B, V1 = c->V1, E = c->E;
  const size_t smem = sizeof(float) * V1;
  const bool pert = c->store_perturbed && c->mode == COOPCAP_SAMPLE_ST_GUMBEL;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(st_bwd_kernel<float, false>), int(smem)))) return rc;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(st_bwd_kernel<bf16, false>), int(smem)))) return rc;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(st_bwd_kernel<float, true>), int(smem)))) return rc;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(st_bwd_kernel<bf16, true>), int(smem)))) return rc;
  // several steps per launch: M = chunk * B rows give properly sized GEMM tiles instead of
  // n_steps launches at the latency floor (g_ws holds `g_chunk_steps` steps)
  const int chunk = demb16 ? (g_chunk_steps < 1 ? 1 : g_chunk_steps) : c->n_steps;
  for (int t0 = 0; t0 < c->n_steps; t0 += chunk) {
    const int nt = min(chunk, c->n_steps - t0);
    const int64_t rows = int64_t(nt) * B;
    const __half* z_t = reinterpret_cast<const __half*>(c->z16_all) + int64_t(t0) * B * V1;
    const float* n_t = c->noise ? c->noise + int64_t(t0) * B * V1 : nullptr;
    bf16* dz_t = reinterpret_cast<bf16*>(dz16) + int64_t(t0) * B * V1;
    if (demb16) {
      // g = demb[t0..] . W_emb^T   ([rows,E] x [V1,E]^T), kept in bf16: half the bytes of this
      // [4096, 9488] intermediate in both directions
      EpiStoreParams e = {};
      e.alpha = 1.f; e.C16 = reinterpret_cast<bf16*>(g_ws); e.ldc16 = V1;
      rc = gemm_run(0, 0, 0, reinterpret_cast<const bf16*>(demb16) + int64_t(t0) * B * E, E, w_emb16,
                    E, int(rows), V1, E, 1, 0, e, s);
      if (rc) return rc;
      if (pert && V1 % 8 == 0 && V1 <= 8 * 256 * ST_REG_IT) {
        CC_CHECK_CUDA(launch_pdl(st_bwd_pert_kernel, dim3((unsigned)rows), dim3(256), size_t(0), s, z_t,
                                 static_cast<const bf16*>(reinterpret_cast<bf16*>(g_ws)), V1, c->inv_tau,
                                 static_cast<const float*>(c->y_max + int64_t(t0) * B),
                                 static_cast<const float*>(c->y_sum + int64_t(t0) * B),
                                 static_cast<const uint8_t*>(c->unfinished + int64_t(t0) * B), dz_t));
      } else
      CC_CHECK_CUDA(launch_pdl(
          pert ? st_bwd_kernel<bf16, true> : st_bwd_kernel<bf16, false>, dim3((unsigned)rows), dim3(256), smem, s, z_t,
          static_cast<const bf16*>(reinterpret_cast<bf16*>(g_ws)), int64_t(V1), V1, c->mode, c->inv_tau, n_t,
          c->seed, uint64_t(SITE_NOISE + t0), B, static_cast<const float*>(c->y_max + int64_t(t0) * B),
          static_cast<const float*>(c->y_sum + int64_t(t0) * B),
          static_cast<const uint8_t*>(c->unfinished + int64_t(t0) * B), dz_t,
          static_cast<const float*>(c->lse + int64_t(t0) * B)));
    } else {
      const float* g_t = reinterpret_cast<const float*>(g_ws) + int64_t(t0) * B * ldg;   // dense, all steps
      CC_CHECK_CUDA(launch_pdl(
          pert ? st_bwd_kernel<float, true> : st_bwd_kernel<float, false>, dim3((unsigned)rows), dim3(256), smem, s, z_t, g_t, ldg, V1, c->mode,
          c->inv_tau, n_t, c->seed, uint64_t(SITE_NOISE + t0), B,
          static_cast<const float*>(c->y_max + int64_t(t0) * B),
          static_cast<const float*>(c->y_sum + int64_t(t0) * B),
          static_cast<const uint8_t*>(c->unfinished + int64_t(t0) * B), dz_t,
          static_cast<const float*>(c->lse + int64_t(t0) * B)));
    }
    // logits + upstream gradient (+ injected noise) read, bf16 dz written
    CC_LAUNCH_CHECK_K(PROF_ST_BWD, s, 0.0,
                      double(rows) * V1 * (2.0 + (demb16 ? 2.0 : 4.0) + 2.0 + ((c->noise && !pert) ? 4.0 : 0.0)));
  }
  return CC_OK;
}

int logp_backward(const coopcap_speaker* c, const int64_t* tok, const float* coef, void* dz16,
                  bool accumulate, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  CC_REQUIRE(!(c->store_perturbed && c->mode == COOPCAP_SAMPLE_ST_GUMBEL),
             "logp_backward: this pass stored perturbed logits (store_perturbed = 1); softmax(z) cannot be rebuilt");
  if (accumulate)
    logp_bwd_kernel<true><<<c->n_steps * c->B, 256, 0, s>>>(reinterpret_cast<const __half*>(c->z16_all), c->V1,
                                                            c->lse, tok, coef, reinterpret_cast<bf16*>(dz16));
  else
    logp_bwd_kernel<false><<<c->n_steps * c->B, 256, 0, s>>>(reinterpret_cast<const __half*>(c->z16_all), c->V1,
                                                             c->lse, tok, coef, reinterpret_cast<bf16*>(dz16));
  CC_LAUNCH_CHECK_K(PROF_LOGP_BWD, s, 0.0, 0.0);
  return CC_OK;
}

int speaker_decode_bwd(const coopcap_speaker* c, const coopcap_speaker_grads* g, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  CC_REQUIRE(g != nullptr && g->dz16 != nullptr, "speaker_decode_bwd: null grads / dz16");
  const bool ps = (c->mode == COOPCAP_SAMPLE_PS_GUMBEL || c->mode == COOPCAP_SAMPLE_PS_MULTINOMIAL);
  if (ps)
    CC_REQUIRE(g->ps_g && g->ps_dpre16 && c->soft16 && c->w_embed16 &&
                   (g->ps_demb16 == nullptr || g->ps_w_emb16 != nullptr),
               "speaker_decode_bwd: partial-sampling pass needs ps_g, ps_dpre16 (+ ps_w_emb16)");
  const int B = c->B, R = c->R, E = c->E, A = c->A, V1 = c->V1, NL = c->NL, n = c->n_steps;
  const int NS = 5 * R + A, XH = E + R;
  const int64_t rows = int64_t(n) * B;
  const bf16* dz16 = reinterpret_cast<const bf16*>(g->dz16);
  bf16* dscat16 = reinterpret_cast<bf16*>(g->dscat16);
  const bf16* xh16 = reinterpret_cast<const bf16*>(c->xh16);
  const bf16* out16 = reinterpret_cast<const bf16*>(c->out16);
  const float scale = c->drop_p > 0.f ? 1.f / (1.f - c->drop_p) : 1.f;

  CC_REQUIRE(g->phase >= 0 && g->phase <= 2 && (!ps || g->phase == 0),
             "speaker_decode_bwd: phase %d (partial-sampling passes take 0 only)", g->phase);
  // logit layer: d_out = dz . W_logit ; g_w_logit = dz^T . out ; g_b_logit = colsum(dz)
  if (!ps && g->phase != 2) {
    EpiStoreParams e = {};
    e.alpha = 1.f; e.C = g->d_out; e.ldc = R;
    if ((rc = gemm_run(0, 0, 1, dz16, V1, c->w_logit16, R, int(rows), R, V1, 1, 0, e, s))) return rc;
    if ((rc = wgrad(dz16, V1, out16, R, V1, R, int(rows), g->g_w_logit, R, s))) return rc;
    if ((rc = colsum_bf16(dz16, rows, V1, V1, g->g_b_logit, s))) return rc;
  }
  if (g->phase == 1) return CC_OK;
  bf16* ps_dpre16 = reinterpret_cast<bf16*>(g->ps_dpre16);
  if (ps) {
    const size_t smem = sizeof(float) * V1;
    if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(st_bwd_kernel<float>), int(smem)))) return rc;
  }

  const size_t att_smem = attention_smem_bytes(A, R, c->L);
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_bwd_kernel), int(att_smem)))) return rc;
  // BPTT.  d[x|h] = dscat . w_cat has only 8 x 8 output tiles for K = 5R+A: it runs split along K
  // with TMA reduce-add into a zeroed buffer so that the whole machine works on it.
  const int dxh_split = (int64_t(B) * XH <= int64_t(128) * 128 * 74 && NS >= 2048) ? 2 : 1;
  // (the slices of d_xh are zeroed step by step inside lstm_bwd_kernel)
  for (int t = n - 1; t >= 0; --t) {
    const float* s_t = c->s_all + int64_t(t) * B * NS;
    bf16* ds_t = dscat16 + int64_t(t) * B * NS;
    float* dc_in = (t == n - 1) ? nullptr : g->dc + int64_t((t + 1) & 1) * B * R;
    float* dc_out = g->dc + int64_t(t & 1) * B * R;
    const float* dh_next = (t == n - 1) ? nullptr : g->d_xh + int64_t(t + 1) * B * XH + E;
    if (ps) {
      // d(loss)/d(v_t) = consumer gradient + d(x_{t+1} pre-activation) . embed^T, then the sampler's
      // softmax backward and the logit dgrad of this step
      const bool dense = (g->ps_demb16 == nullptr);
      const int64_t ldg = dense ? g->ps_ldg : V1;
      float* g_t = dense ? g->ps_g + int64_t(t) * B * ldg : g->ps_g;
      if (!dense) {
        EpiStoreParams e = {};
        e.alpha = 1.f; e.C = g_t; e.ldc = ldg;
        if ((rc = gemm_run(0, 0, 0, reinterpret_cast<const bf16*>(g->ps_demb16) + int64_t(t) * B * E, E,
                           g->ps_w_emb16, E, B, V1, E, 1, 0, e, s)))
          return rc;
      }
      if (t + 1 < n) {
        EpiStoreParams e = {};
        e.alpha = 1.f; e.C = g_t; e.ldc = ldg; e.mode = 1;
        if ((rc = gemm_run(0, 0, 0, ps_dpre16 + int64_t(t + 1) * B * E, E, c->w_embed16, E, B, V1, E,
                           1, 0, e, s)))
          return rc;
      }
      bf16* dz_t = const_cast<bf16*>(dz16) + int64_t(t) * B * V1;
      CC_CHECK_CUDA(launch_pdl(
          st_bwd_kernel<float>, dim3((unsigned)B), dim3(256), sizeof(float) * V1, s,
          reinterpret_cast<const __half*>(c->z16_all) + int64_t(t) * B * V1, static_cast<const float*>(g_t),
          ldg, V1, c->mode, c->inv_tau,
          c->noise ? c->noise + int64_t(t) * B * V1 : static_cast<const float*>(nullptr), c->seed,
          uint64_t(SITE_NOISE + t), B, static_cast<const float*>(c->y_max + int64_t(t) * B),
          static_cast<const float*>(c->y_sum + int64_t(t) * B),
          static_cast<const uint8_t*>(c->unfinished + int64_t(t) * B), dz_t,
          static_cast<const float*>(c->lse + int64_t(t) * B)));
      CC_LAUNCH_CHECK_K(PROF_ST_BWD, s, 0.0, double(B) * V1 * 8.0);
      EpiStoreParams e = {};
      e.alpha = 1.f; e.C = g->d_out + int64_t(t) * B * R; e.ldc = R;
      if ((rc = gemm_run(0, 0, 1, dz_t, V1, c->w_logit16, R, B, R, V1, 1, 0, e, s))) return rc;
    }
    {
      const int nthr = B * (R / 4);
      CC_CHECK_CUDA(launch_pdl(
          lstm_bwd_kernel, dim3((nthr + 255) / 256), dim3(256), 0, s, g->d_out + int64_t(t) * B * R,
          out16 + int64_t(t) * B * R, dh_next, int64_t(XH), dc_in, dc_out, s_t, int64_t(NS),
          c->u_all + int64_t(t) * B * 2 * R, c->c_all + int64_t(t) * B * R,
          c->c_all + int64_t(t + 1) * B * R, ds_t, c->drop_p, B, R,
          (dxh_split > 1 && XH % 4 == 0) ? g->d_xh + int64_t(t) * B * XH : static_cast<float*>(nullptr), XH));
      CC_LAUNCH_CHECK_K(PROF_LSTM, s, 0.0, 0.0);
    }
    // d_att_res = d_u . W_a2c      ([B,2R] x [2R,R])
    float* dres_t = g->d_att_res + int64_t(t) * B * R;
    {
      EpiStoreParams e = {};
      e.alpha = 1.f; e.C = dres_t; e.ldc = R;
      if ((rc = gemm_run(0, 0, 1, ds_t + 3 * R, NS, c->w_a2c16, R, B, R, 2 * R, 1, 0, e, s))) return rc;
    }
    static const bool att_one_pass = [] { const char* e = getenv("COOPCAP_ATT_BWD_TWO_PASS"); return !(e && e[0] == '1'); }();
    if (A == 512 && R == 512 && c->att_res32 && att_one_pass) {
      // single pass over (att_e, p_att): the softmax backward's mean comes from the saved fp32 att_res
      if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_bwd5_kernel<512>), ATT4_SMEM))) return rc;
      CC_CHECK_CUDA(launch_pdl(attention_bwd5_kernel<512>, dim3(std::min(num_sms(), B)),
                               dim3(ATT4_THREADS), size_t(ATT4_SMEM), s,
                               reinterpret_cast<const bf16*>(c->p_att16),
                               reinterpret_cast<const bf16*>(c->att_e16), c->att_off, c->L,
                               c->att_order, s_t, int64_t(NS), 5 * R, c->w_alpha, dres_t,
                               c->att_res32 + int64_t(t) * B * R, c->att_w + int64_t(t) * NL,
                               g->de + int64_t(t) * NL, ds_t, B));
    } else if (A == 512 && R == 512) {
      if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_bwd4_kernel<512>), ATT4_SMEM))) return rc;
      CC_CHECK_CUDA(launch_pdl(attention_bwd4_kernel<512>, dim3(std::min(num_sms(), B)),
                               dim3(ATT4_THREADS), size_t(ATT4_SMEM), s,
                               reinterpret_cast<const bf16*>(c->p_att16),
                               reinterpret_cast<const bf16*>(c->att_e16), c->att_off, c->L,
                               c->att_order, s_t,
                               int64_t(NS), 5 * R, c->w_alpha, dres_t, c->att_w + int64_t(t) * NL,
                               g->de + int64_t(t) * NL, ds_t, B));
    } else {
      attention_bwd_kernel<<<B, ATT_THREADS, att_smem, s>>>(
          reinterpret_cast<const bf16*>(c->p_att16), reinterpret_cast<const bf16*>(c->att_e16),
          c->att_off, c->L, s_t, NS, 5 * R, c->w_alpha, dres_t, c->att_w + int64_t(t) * NL,
          g->de + int64_t(t) * NL, ds_t, A, R);
    }
    CC_LAUNCH_CHECK_K(PROF_ATT_BWD, s, 0.0,
                      2.0 * NL * (A + R) + 4.0 * B * (A + R) + 8.0 * NL + 2.0 * B * A);
    // d[x_t | h_{t-1}] = dscat . w_cat      ([B,5R+A] x [5R+A,E+R])
    {
      EpiStoreParams e = {};
      e.alpha = 1.f; e.C = g->d_xh + int64_t(t) * B * XH; e.ldc = XH;
      e.mode = dxh_split > 1 ? 2 : 0;
      if ((rc = gemm_run(0, 0, 1, ds_t, NS, c->w_cat16, XH, B, XH, NS, dxh_split,
                         dxh_split > 1 ? 128 : 0, e, s)))
        return rc;
    }
    if (ps && t > 0) {
      ps_dpre_kernel<<<B, 128, 0, s>>>(g->d_xh + int64_t(t) * B * XH, xh16 + int64_t(t) * B * XH, E, XH,
                                       scale, ps_dpre16 + int64_t(t) * B * E);
      CC_LAUNCH_CHECK_K(PROF_REDUCE, s, 0.0, 0.0);
    }
  }
  if (ps) {
    if ((rc = wgrad(dz16, V1, out16, R, V1, R, int(rows), g->g_w_logit, R, s))) return rc;
    if ((rc = colsum_bf16(dz16, rows, V1, V1, g->g_b_logit, s))) return rc;
  }
  // step-batched weight gradients
  if ((rc = wgrad(dscat16, NS, xh16, XH, 5 * R, E, int(rows), g->g_w_i2h, E, s))) return rc;
  if ((rc = wgrad(dscat16, NS, xh16 + E, XH, 5 * R, R, int(rows), g->g_w_h2h, R, s))) return rc;
  if ((rc = wgrad(dscat16 + 5 * R, NS, xh16 + E, XH, A, R, int(rows), g->g_w_h2att, R, s))) return rc;
  if ((rc = wgrad(dscat16 + 3 * R, NS, c->att_res16, R, 2 * R, R, int(rows), g->g_w_a2c, R, s))) return rc;
  if ((rc = colsum_bf16(dscat16, rows, 5 * R, NS, g->g_b_gates, s))) return rc;
  if ((rc = colsum_bf16(dscat16 + 5 * R, rows, A, NS, g->g_b_h2att, s))) return rc;
  // the a2c bias enters the two maxout pre-activations (columns 3R..5R) like the gate bias does: its
  // gradient is that slice of g_b_gates (was a third pass over 34 MB of dscat16)
  CC_CHECK_CUDA(cudaMemcpyAsync(g->g_b_a2c, g->g_b_gates + 3 * R, sizeof(float) * 2 * R,
                                cudaMemcpyDeviceToDevice, s));
  // input embedding
  if (ps) {
    // steps >= 1 were fed v_{t-1} . embed:  g_embed[:V1] = sum_t v_{t-1}^T . dpre_t ; step 0 is a lookup
    if (n > 1 &&
        (rc = wgrad(c->soft16, V1, ps_dpre16 + int64_t(B) * E, E, V1, E, int(rows - B), g->g_embed, E, s)))
      return rc;
    CC_CHECK_CUDA(launch_pdl(embed_grad_kernel, dim3((unsigned)B), dim3(128), size_t(0), s, c->tok_fed, g->d_xh, xh16, E, XH, scale, g->g_embed));
  } else {
    CC_CHECK_CUDA(launch_pdl(embed_grad_kernel, dim3((unsigned)rows), dim3(128), size_t(0), s, c->tok_fed, g->d_xh, xh16, E, XH, scale, g->g_embed));
  }
  CC_LAUNCH_CHECK_K(PROF_REDUCE, s, 0.0, 0.0);
  // region tensors: deferred accumulation over the steps, then the prologue layers
  {
    // galpha / bias partials live in d_out (free after the BPTT loop): 2 x [B, A] <= [cap*B, R]
    float* galpha_part = g->d_out;
    float* gbias_part = g->d_out + int64_t(B) * A;
    if (A == 512 && R == 512) {
      constexpr int ST = 3;
      const size_t smem = attention_deferred2_smem<512, ST>(c->L, n);
      CC_REQUIRE(smem <= 227 * 1024, "deferred attention backward needs %zu B of shared memory", smem);
      if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_deferred2_kernel<512, ST>), int(smem)))) return rc;
      CC_CHECK_CUDA(launch_pdl(attention_deferred2_kernel<512, ST>, dim3(B), dim3(ATT_THREADS), size_t(smem), s, 
          reinterpret_cast<const bf16*>(c->p_att16), c->att_off, c->L, c->s_all, NS, 5 * R,
          int64_t(B) * NS, c->w_alpha, g->d_att_res, c->att_w, g->de, NL, n, B, g->d_att_e,
          reinterpret_cast<bf16*>(g->d_p_att16), galpha_part, gbias_part, c->att_order));
    } else {
      const int Lp = (c->L + 3) & ~3;
      const int groups = ATT_THREADS / (A / 8);
      const size_t smem = sizeof(float) * (size_t(n) * (A + R) + 2 * size_t(n) * Lp + size_t(groups) * A);
      CC_REQUIRE(smem <= 227 * 1024, "deferred attention backward needs %zu B of shared memory", smem);
      if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_deferred_bwd_kernel), int(smem)))) return rc;
      attention_deferred_bwd_kernel<<<B, ATT_THREADS, smem, s>>>(
          reinterpret_cast<const bf16*>(c->p_att16), c->att_off, c->L, c->s_all, NS, 5 * R,
          int64_t(B) * NS, c->w_alpha, g->d_att_res, c->att_w, g->de, NL, n, B, g->d_att_e,
          reinterpret_cast<bf16*>(g->d_p_att16), galpha_part, gbias_part, A, R);
    }
    // p_att read, d_p_att (bf16) + d_att_e (fp32) written, per-step vectors read
    CC_LAUNCH_CHECK_K(PROF_ATT_DEFERRED, s, 0.0,
                      2.0 * NL * A + 2.0 * NL * A + 4.0 * NL * R +
                          double(n) * (4.0 * B * (A + R) + 8.0 * NL));
    if ((rc = colsum_f32(galpha_part, B, A, A, g->g_w_alpha, s))) return rc;
    if ((rc = colsum_f32(gbias_part, B, A, A, g->g_b_ctx2att, s))) return rc;
  }
  {
    // d_att_e += d_p_att . W_ctx2att     ([NL,A] x [A,R])
    EpiStoreParams e = {};
    e.alpha = 1.f; e.C = g->d_att_e; e.ldc = R; e.mode = 1;
    if ((rc = gemm_run(0, 0, 1, g->d_p_att16, A, c->w_ctx2att16, R, NL, R, A, 1, 0, e, s))) return rc;
    const int64_t n4 = int64_t(NL) * R / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > int64_t(num_sms()) * 16) blocks = int64_t(num_sms()) * 16;
    CC_CHECK_CUDA(launch_pdl(mask_pre_kernel, dim3((unsigned)blocks), dim3(256), size_t(0), s, g->d_att_e,
                                                     reinterpret_cast<const bf16*>(c->att_e16), n4,
                                                     scale, reinterpret_cast<bf16*>(g->d_pre16)));
    CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 0.0);
    if ((rc = wgrad(g->d_p_att16, A, c->att_e16, R, A, R, NL, g->g_w_ctx2att, R, s))) return rc;
    if ((rc = wgrad(g->d_pre16, R, c->att16, c->D, R, c->D, NL, g->g_w_att_embed, c->D, s))) return rc;
    if ((rc = colsum_bf16(g->d_pre16, NL, R, R, g->g_b_att_embed, s))) return rc;
  }
  return CC_OK;
}

// flat fp32 bucket: g *= grad_scale; clamp; Adam  (optimizer.py:233-242, misc/utils.py:65-69)
__global__ void clamp_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                  float* __restrict__ m, float* __restrict__ v, int64_t n,
                                  float grad_scale, float clip, float step_size, float b1c, float b2,
                                  float b2c, float eps, float wd, float bc2_sqrt) {
  pdl_launch_dependents();
  pdl_wait();
  // b1c = 1 - beta1, b2c = 1 - beta2, rounded from double like torch's scalar arguments
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4;
       i += int64_t(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float x = gg[q] * grad_scale;
      if (clip > 0.f) x = fminf(fmaxf(x, -clip), clip);
      x += wd * pp[q];
      mm[q] = fmaf(b1c, x - mm[q], mm[q]);           // torch: exp_avg.lerp_(grad, 1 - beta1)
      vq[q] = fmaf(b2c * x, x, b2 * vq[q]);          // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
      const float denom = sqrtf(vq[q]) / bc2_sqrt + eps;
      pp[q] -= step_size * (mm[q] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vq[0], vq[1], vq[2], vq[3]);
  }
  // tail
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float x = g[i] * grad_scale;
      if (clip > 0.f) x = fminf(fmaxf(x, -clip), clip);
      x += wd * p[i];
      const float mq = fmaf(b1c, x - m[i], m[i]);
      const float vq = fmaf(b2c * x, x, b2 * v[i]);
      m[i] = mq; v[i] = vq;
      p[i] -= step_size * (mq / (sqrtf(vq) / bc2_sqrt + eps));
    }
  }
}

}  // namespace coopcap

extern "C" {

int coopcap_st_backward(const coopcap_speaker* ctx, const void* demb16, const void* w_emb16,
                        void* g_ws, int g_chunk_steps, void* dz16, coopcap_stream_t stream) {
  if (!demb16 || !w_emb16 || !g_ws) {
    coopcap::set_last_error("st_backward: null demb16 / w_emb16 / g_ws");
    return coopcap::CC_ERR_ARG;
  }
  return coopcap::st_backward(ctx, demb16, w_emb16, g_ws, g_chunk_steps, ctx ? ctx->V1 : 0, dz16,
                              reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_st_backward_dense(const coopcap_speaker* ctx, const float* g, int64_t ldg, void* dz16,
                              coopcap_stream_t stream) {
  if (!g) {
    coopcap::set_last_error("st_backward_dense: null g");
    return coopcap::CC_ERR_ARG;
  }
  return coopcap::st_backward(ctx, nullptr, nullptr, const_cast<float*>(g), 0, ldg, dz16,
                              reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_logp_backward(const coopcap_speaker* ctx, const int64_t* tok, const float* coef,
                          void* dz16, coopcap_stream_t stream) {
  return coopcap::logp_backward(ctx, tok, coef, dz16, false, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_logp_backward_acc(const coopcap_speaker* ctx, const int64_t* tok, const float* coef,
                              void* dz16, coopcap_stream_t stream) {
  return coopcap::logp_backward(ctx, tok, coef, dz16, true, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_speaker_decode_bwd(const coopcap_speaker* ctx, const coopcap_speaker_grads* gr,
                               coopcap_stream_t stream) {
  return coopcap::speaker_decode_bwd(ctx, gr, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_clamp_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       double grad_scale, double clip, double lr, double beta1, double beta2,
                       double eps, double weight_decay, int step, coopcap_stream_t stream) {
  using namespace coopcap;
  if (n <= 0) return CC_OK;
  CC_REQUIRE(step >= 1, "clamp_adam: step must be >= 1");
  CC_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
               reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
             "clamp_adam: buffers must be 16-byte aligned");
  // bias corrections in double, as torch.optim.Adam computes them on the host (1 - beta ** step)
  const double bc1 = 1.0 - pow(beta1, double(step));
  const double bc2 = 1.0 - pow(beta2, double(step));
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > int64_t(num_sms()) * 8) blocks = int64_t(num_sms()) * 8;
  if (blocks < 1) blocks = 1;
  CC_CHECK_CUDA(launch_pdl(clamp_adam_kernel, dim3((unsigned)blocks), dim3(256), size_t(0), reinterpret_cast<cudaStream_t>(stream), 
      param, grad, exp_avg, exp_avg_sq, n, float(grad_scale), float(clip), float(lr / bc1),
      float(1.0 - beta1), float(beta2), float(1.0 - beta2), float(eps), float(weight_decay),
      float(sqrt(bc2))));
  CC_LAUNCH_CHECK_K(PROF_ADAM, reinterpret_cast<cudaStream_t>(stream), 0.0, 28.0 * double(n));
  return CC_OK;
}

}  // extern "C"
