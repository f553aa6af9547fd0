// Speaker (Att2in2) forward: weight packing, prologue (att_embed / ctx2att over packed regions),
// and the decode loop (gate GEMM -> additive attention -> a2c GEMM -> maxout-LSTM pointwise ->
// logit GEMM with the sampler in its epilogue -> per-row finish + next-input gather).  See include/coopcap.h for the buffer layout and
// the reference lines each piece replaces.
#include <algorithm>
#include <atomic>
#include "../../include/coopcap.h"
#include "common.cuh"
#include "gemm.cuh"
#include "speaker_kernels.cuh"
#include "cell_step.cuh"
#include "attention.cuh"
#include "logit_sample.cuh"

namespace coopcap {

using bf16 = __nv_bfloat16;

// ------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------
// fp32 [rows, cols] -> bf16 block of a wider matrix (ld_dst), 4 elements per thread
__global__ void cast_block_kernel(const float* __restrict__ src, int64_t rows, int cols,
                                  bf16* __restrict__ dst, int64_t ld_dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = rows * (cols / 4);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / (cols / 4);
    const int c = int(i % (cols / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * cols + c);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = o;
  }
}

int cast_block(const float* src, int64_t rows, int cols, void* dst, int64_t ld_dst,
               cudaStream_t s) {
  CC_REQUIRE(cols % 4 == 0 && ld_dst % 4 == 0, "cast_block: cols %d / ld %lld not multiple of 4",
             cols, (long long)ld_dst);
  if (rows <= 0) return CC_OK;
  const int64_t n4 = rows * (cols / 4);
  int64_t want = (n4 + 255) / 256, cap_blocks = int64_t(num_sms()) * 16;
  int grid = int(want < cap_blocks ? want : cap_blocks);
  CC_CHECK_CUDA(launch_pdl(cast_block_kernel, dim3(grid), dim3(256), size_t(0), s, src, rows, cols, reinterpret_cast<bf16*>(dst), ld_dst));
  CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 0.0);
  return CC_OK;
}

__global__ void bias_cat_kernel(const float* __restrict__ b_i2h, const float* __restrict__ b_h2h,
                                const float* __restrict__ b_h2att, int n5r, int A,
                                float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n5r) out[i] = b_i2h[i] + b_h2h[i];
  else if (i < n5r + A) out[i] = b_h2att[i - n5r];
}

// Several fp32 -> bf16 block casts in ONE launch (the weight re-pack after every optimizer step is
// eleven small casts: as separate launches their latency, not their 157 MB, set the cost).
// Sources are dense [rows, cols] fp32 with cols % 8 == 0; destinations bf16 with row pitch ld_dst.
constexpr int CAST_MULTI_MAX = 12;
struct CastJobs {
  const float* src[CAST_MULTI_MAX];
  bf16* dst[CAST_MULTI_MAX];
  int cols[CAST_MULTI_MAX];
  int64_t ld_dst[CAST_MULTI_MAX];
  int64_t first[CAST_MULTI_MAX + 1];      // prefix sums of the jobs' 8-element groups
  int n;
};
__global__ void __launch_bounds__(256) cast_multi_kernel(const CastJobs J) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = J.first[J.n];
  for (int64_t g = int64_t(blockIdx.x) * 256 + threadIdx.x; g < total; g += int64_t(gridDim.x) * 256) {
    int k = 0;
    while (k + 1 < J.n && g >= J.first[k + 1]) ++k;
    const int64_t e = (g - J.first[k]) * 8;            // flat element index inside job k
    const int cols = J.cols[k];
    const int64_t row = e / cols;
    const int col = int(e - row * cols);
    const float4 a = __ldcs(reinterpret_cast<const float4*>(J.src[k] + e));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(J.src[k] + e + 4));
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<uint4*>(J.dst[k] + row * J.ld_dst[k] + col) = float8_to_bf16x8(f);
  }
}
struct CastBatch {
  CastJobs J = {};
  bool ok = true;
  void add(const float* src, int64_t rows, int cols, void* dst, int64_t ld_dst) {
    if (J.n >= CAST_MULTI_MAX || cols % 8 != 0 || ld_dst % 8 != 0 ||
        (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
      ok = false;
      return;
    }
    const int k = J.n++;
    J.src[k] = src; J.dst[k] = reinterpret_cast<bf16*>(dst); J.cols[k] = cols; J.ld_dst[k] = ld_dst;
    J.first[k + 1] = J.first[k] + rows * cols / 8;
  }
  int run(cudaStream_t s) {
    if (J.n == 0) return CC_OK;
    const int64_t groups = J.first[J.n];
    int64_t grid = (groups + 255) / 256;
    static const int per_sm = [] { const char* e = getenv("COOPCAP_CAST_CTAS_PER_SM"); return e ? atoi(e) : 16; }();
    const int64_t cap = int64_t(num_sms()) * per_sm;
    if (grid > cap) grid = cap;
    CC_CHECK_CUDA(launch_pdl(cast_multi_kernel, dim3(unsigned(grid)), dim3(256), size_t(0), s, J));
    CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 0.0);
    return CC_OK;
  }
};
static std::atomic<int> g_cast_multi{[] { const char* e = getenv("COOPCAP_CAST_MULTI"); return (e && e[0] == '0') ? 0 : 1; }()};
void set_cast_multi(int on) { g_cast_multi.store(on ? 1 : 0, std::memory_order_relaxed); }
int cast_blocks(int n, const float* const* src, const int64_t* rows, const int* cols, void* const* dst,
                const int64_t* ld_dst, cudaStream_t s) {
  CastBatch b;
  for (int i = 0; i < n; ++i) b.add(src[i], rows[i], cols[i], dst[i], ld_dst[i]);
  // One fused launch for the eleven casts of a weight re-pack unless the host side asked for separate
  // launches (coopcap_set_cast_multi(0) / COOPCAP_CAST_MULTI=0).  Measured A/B on one box (r2,
  // gpurun_out/s7_bench*.json): fused shortens the device-resident step and the feature-store
  // end-to-end step by 0.04 ms (5.583 -> 5.541, 5.568 -> 5.547 ms), but lengthens the end-to-end step
  // fed by the ZERO-COPY upload kernel -- the next batch's reader shares the SMs at that point -- from
  // 10.57 to 11.79 ms, so data.upload_batch(zero_copy=True) switches it off.
  const bool fused = g_cast_multi.load(std::memory_order_relaxed) != 0;
  if (b.ok && fused) return b.run(s);
  int rc;
  for (int i = 0; i < n; ++i)
    if ((rc = cast_block(src[i], rows[i], cols[i], dst[i], ld_dst[i], s))) return rc;
  return CC_OK;
}

// ------------------------------------------------------------------------------------------
// prologue: cast + pack valid regions
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_row(const int* __restrict__ off, int B, int r) {
  // largest b with off[b] <= r
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= r) lo = mid; else hi = mid;
  }
  return lo;
}

// Cast + pack: packed row r (valid region l of batch row b) <- att[b, l, :].  Work item =
// PACK_ROWS consecutive packed rows; every thread first issues all of its 16-byte loads for the
// item (4 per row at D = 2048 -- this kernel also runs as the zero-copy PCIe reader), then
// converts and stores.  128 threads / <= 64 registers per CTA (PACK_ROWS <= 2) so that a resident
// pack CTA still leaves room for a 320-thread GEMM CTA on the same SM (the upload overlaps compute).
constexpr int PACK_THREADS = 128;
template <int PACK_ROWS>
__global__ void __launch_bounds__(PACK_THREADS)
pack_att_kernel(const float* __restrict__ att, const int* __restrict__ off, int B, int L, int D,
                int NL, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int d4 = D / 4;                      // float4 per row
  const int n_items = (NL + PACK_ROWS - 1) / PACK_ROWS;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int r0 = item * PACK_ROWS;
    const float4* src[PACK_ROWS];
#pragma unroll
    for (int k = 0; k < PACK_ROWS; ++k) {
      const int r = min(r0 + k, NL - 1);
      int64_t src_row = r;
      if (off) {
        const int b = find_row(off, B, r);
        src_row = int64_t(b) * L + (r - off[b]);
      }
      src[k] = reinterpret_cast<const float4*>(att + src_row * D);
    }
    for (int i0 = 0; i0 < d4; i0 += 4 * PACK_THREADS) {   // 4 float4 per thread per row
      float4 v[PACK_ROWS][4];
#pragma unroll
      for (int k = 0; k < PACK_ROWS; ++k)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int i = i0 + h * PACK_THREADS + threadIdx.x;
          v[k][h] = (i < d4) ? __ldcs(src[k] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int k = 0; k < PACK_ROWS; ++k) {
        if (r0 + k >= NL) break;
        uint2* dst = reinterpret_cast<uint2*>(out + int64_t(r0 + k) * D);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int i = i0 + h * PACK_THREADS + threadIdx.x;
          if (i < d4) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v[k][h].x, v[k][h].y);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(v[k][h].z, v[k][h].w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&a);
            o.y = *reinterpret_cast<uint32_t*>(&b2);
            dst[i] = o;
          }
        }
      }
    }
  }
}

// step 0: feed the start token, zero h_{-1} and c_{-1}
__global__ void start_step_kernel(const float* __restrict__ embed, int64_t start_scalar,
                                  const int64_t* __restrict__ start_rows, int B, int E, int R,
                                  const uint8_t* __restrict__ keep_embed, uint64_t seed,
                                  float drop_p, bf16* __restrict__ xh0, float* __restrict__ c0,
                                  int64_t* __restrict__ tok_fed0) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int64_t start = start_rows ? start_rows[b] : start_scalar;
  bf16* row = xh0 + int64_t(b) * (E + R);
  embed_row(embed, start, E, keep_embed ? keep_embed + int64_t(b) * E : nullptr, seed,
            SITE_DROP_EMBED, int64_t(b) * E, drop_p, row);
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    row[E + i] = __float2bfloat16_rn(0.f);
    c0[int64_t(b) * R + i] = 0.f;
  }
  if (threadIdx.x == 0) tok_fed0[b] = start;
}

// ------------------------------------------------------------------------------------------
// additive attention forward                                           (AttModel.py:465-489)
//   e_l = sum_j alpha_j tanh(p_att[l,j] + att_h[j]);  w = softmax_l(e) over the valid regions
//   att_res[j] = sum_l w_l att_e[l,j]
// one CTA per batch row; HBM-bound: reads p_att[b] and att_e[b] once (bf16, 16-byte loads).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ATT_THREADS)
attention_fwd_kernel(const bf16* __restrict__ p_att16, const bf16* __restrict__ att_e16,
                     const int* __restrict__ off, int Lfix, const float* __restrict__ s_row0,
                     int64_t lds, int att_h_col, const float* __restrict__ w_alpha,
                     bf16* __restrict__ att_res16, float* __restrict__ att_w, int A, int R) {
  extern __shared__ float sm[];
  float* s_ah = sm;              // [A]
  float* s_al = sm + A;          // [A]
  float* s_red = sm + 2 * A;     // [8]
  float* s_e = sm + 2 * A + 8;   // [Lb]
  const int b = blockIdx.x;
  const int r0 = off ? off[b] : b * Lfix;
  const int Lb = off ? off[b + 1] - r0 : Lfix;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
  for (int i = threadIdx.x; i < A; i += ATT_THREADS) {
    s_ah[i] = att_h[i];
    s_al[i] = w_alpha[i];
  }
  __syncthreads();
  // phase 1: scores, one warp per region
  for (int l = warp; l < Lb; l += ATT_THREADS / 32) {
    const uint4* prow = reinterpret_cast<const uint4*>(p_att16 + int64_t(r0 + l) * A);
    float acc = 0.f;
    for (int c = lane; c < A / 8; c += 32) {
      const uint4 u = __ldg(prow + c);
      float f[8];
      bf16x8_to_float(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += s_al[c * 8 + j] * tanh_fast(f[j] + s_ah[c * 8 + j]);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_e[l] = acc;
  }
  __syncthreads();
  // phase 2: softmax over the valid regions
  float m = -INFINITY;
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) m = fmaxf(m, s_e[l]);
  m = block_max_256(m, s_red);
  float sum = 0.f;
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) {
    const float e = __expf(s_e[l] - m);
    s_e[l] = e;
    sum += e;
  }
  sum = block_sum_256(sum, s_red);
  const float inv = 1.f / sum;
  __syncthreads();
  for (int l = threadIdx.x; l < Lb; l += ATT_THREADS) {
    const float w = s_e[l] * inv;
    s_e[l] = w;
    att_w[r0 + l] = w;
  }
  __syncthreads();
  // phase 3: weighted sum; thread owns 8 columns, thread groups stride over the regions
  const int tpr = R / 8;               // threads per region row
  const int groups = ATT_THREADS / tpr;
  const int g = threadIdx.x / tpr, c = threadIdx.x % tpr;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < groups) {
    for (int l = g; l < Lb; l += groups) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(att_e16 + int64_t(r0 + l) * R) + c);
      float f[8];
      bf16x8_to_float(u, f);
      const float w = s_e[l];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += w * f[j];
    }
  }
  __syncthreads();
  // cross-group reduction through shared memory (reuses s_ah/s_al: 2A >= groups*R is not
  // guaranteed, so use a dedicated region after s_e)
  float* s_acc = s_e + ((Lb + 3) & ~3);  // [groups][R]
  if (g < groups) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[g * R + c * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < R; j += ATT_THREADS) {
    float t = 0.f;
    for (int gg = 0; gg < groups; ++gg) t += s_acc[gg * R + j];
    att_res16[int64_t(b) * R + j] = __float2bfloat16_rn(t);
  }
}

// ------------------------------------------------------------------------------------------
// maxout-LSTM pointwise forward                                         (AttModel.py:515-530)
// ------------------------------------------------------------------------------------------
__global__ void lstm_fwd_kernel(const float* __restrict__ s, int64_t lds, const float* __restrict__ u,
                                const float* __restrict__ c_prev, float* __restrict__ c_next,
                                bf16* __restrict__ h_dst, int64_t ld_h, bf16* __restrict__ out16,
                                const uint8_t* __restrict__ keep, uint64_t seed, uint64_t stream,
                                float drop_p, int B, int R) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // 4 units per thread
  const int per_row = R / 4;
  if (idx >= B * per_row) return;
  const int b = idx / per_row, j = (idx % per_row) * 4;
  const float* sr = s + int64_t(b) * lds;
  const float4 si = *reinterpret_cast<const float4*>(sr + j);
  const float4 sf = *reinterpret_cast<const float4*>(sr + R + j);
  const float4 so = *reinterpret_cast<const float4*>(sr + 2 * R + j);
  const float4 s1 = *reinterpret_cast<const float4*>(sr + 3 * R + j);
  const float4 s2 = *reinterpret_cast<const float4*>(sr + 4 * R + j);
  const float4 u1 = *reinterpret_cast<const float4*>(u + int64_t(b) * 2 * R + j);
  const float4 u2 = *reinterpret_cast<const float4*>(u + int64_t(b) * 2 * R + R + j);
  const float4 cp = *reinterpret_cast<const float4*>(c_prev + int64_t(b) * R + j);
  const float ai[4] = {si.x, si.y, si.z, si.w}, af[4] = {sf.x, sf.y, sf.z, sf.w};
  const float ao[4] = {so.x, so.y, so.z, so.w};
  const float a1[4] = {s1.x + u1.x, s1.y + u1.y, s1.z + u1.z, s1.w + u1.w};
  const float a2[4] = {s2.x + u2.x, s2.y + u2.y, s2.z + u2.z, s2.w + u2.w};
  const float ac[4] = {cp.x, cp.y, cp.z, cp.w};
  float cn[4], h[4], o[4];
  bool k[4] = {true, true, true, true};
  const float sc = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  if (drop_p > 0.f)
    keep4(keep, (int64_t(b) * R + j) >> 2, seed, stream, drop_p, k);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float ig = 1.f / (1.f + expf(-ai[q]));
    const float fg = 1.f / (1.f + expf(-af[q]));
    const float og = 1.f / (1.f + expf(-ao[q]));
    const float g = fmaxf(a1[q], a2[q]);
    cn[q] = fg * ac[q] + ig * g;
    h[q] = og * tanhf(cn[q]);
    o[q] = k[q] ? h[q] * sc : 0.f;
  }
  *reinterpret_cast<float4*>(c_next + int64_t(b) * R + j) = make_float4(cn[0], cn[1], cn[2], cn[3]);
  {
    __nv_bfloat162 a = __floats2bfloat162_rn(h[0], h[1]), b2 = __floats2bfloat162_rn(h[2], h[3]);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&a);
    w.y = *reinterpret_cast<uint32_t*>(&b2);
    *reinterpret_cast<uint2*>(h_dst + int64_t(b) * ld_h + j) = w;
  }
  {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b2 = __floats2bfloat162_rn(o[2], o[3]);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&a);
    w.y = *reinterpret_cast<uint32_t*>(&b2);
    *reinterpret_cast<uint2*>(out16 + int64_t(b) * R + j) = w;
  }
}

// Partial-sampling modes: the vector a step emits (gumbel_softmax.py:28-40, multinomial_soft.py:21-33).
// One CTA per row; y is rebuilt from the logits, the regenerated / injected noise and the saved
// (max, sum) exactly like the straight-through backward does.  Rows with part_u < ps_prob emit
// one_hot(id), the others y.
__global__ void __launch_bounds__(256)
ps_vec_kernel(const __half* __restrict__ z, int V1, int mode, float inv_tau,
              const float* __restrict__ noise, uint64_t seed, uint64_t nstream,
              const float* __restrict__ ymax, const float* __restrict__ ysum,
              const int64_t* __restrict__ tok_fed_next, float ps_prob,
              const float* __restrict__ part_u, uint64_t pstream, bf16* __restrict__ soft16,
              uint8_t* __restrict__ ps_sel) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const __half* zr = z + int64_t(b) * V1;
  const float* nr = noise ? noise + int64_t(b) * V1 : nullptr;
  bf16* dr = soft16 + int64_t(b) * V1;
  const bool fast = (noise == nullptr);
  const int nv4 = V1 / 4;
  bool hard = false;
  if (ps_prob > 0.f) {
    const float u = part_u ? part_u[b] : Philox::u01(Philox::gen(seed, pstream, uint64_t(b)).x);
    hard = u < ps_prob;
  }
  if (threadIdx.x == 0) ps_sel[b] = hard ? 1 : 0;
  if (hard) {
    const int id = int(tok_fed_next[b]);
    for (int v4 = threadIdx.x; v4 < nv4; v4 += 256)
      store_bf16x4(dr + 4 * v4, 4 * v4 == id ? 1.f : 0.f, 4 * v4 + 1 == id ? 1.f : 0.f,
                   4 * v4 + 2 == id ? 1.f : 0.f, 4 * v4 + 3 == id ? 1.f : 0.f);
    return;
  }
  const float m = ymax[b], inv_s = 1.f / ysum[b];
  for (int v4 = threadIdx.x; v4 < nv4; v4 += 256) {
    float x4[4];
    f16x4_to_float(zr + 4 * v4, x4);
    float u4[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == COOPCAP_SAMPLE_PS_GUMBEL) noise4(nr, v4, seed, nstream, uint64_t(b) * nv4 + v4, u4);
    float y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      y[q] = ex2_ftz((st_score(mode, x4[q], u4[q], inv_tau, fast) - m) * 1.4426950408889634f) * inv_s;
    store_bf16x4(dr + 4 * v4, y[0], y[1], y[2], y[3]);
  }
}

// finished rows emit the EOS one-hot (AttModel.py:428-432); runs after the next-input GEMM has
// consumed the unmasked vectors
__global__ void __launch_bounds__(256)
ps_mask_kernel(const uint8_t* __restrict__ unf, int V1, bf16* __restrict__ soft16) {
  const int b = blockIdx.x;
  if (unf[b]) return;
  bf16* dr = soft16 + int64_t(b) * V1;
  for (int v4 = threadIdx.x; v4 < V1 / 4; v4 += 256)
    store_bf16x4(dr + 4 * v4, v4 == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
}

// caption summary after the last step: k_b = number of leading non-zero ids, n = max_b k_b,
// cap_len[b] = min(k_b + 2, n + 1)                    (AlternatingJointModel.py:353-355)
__global__ void caption_summary_kernel(const int64_t* __restrict__ tok_out, int B, int n_steps,
                                       int* __restrict__ n_out, int* __restrict__ cap_len) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    // first step with a non-positive id; the loads are independent (no early exit), so they overlap
    int k = n_steps;
    for (int t = n_steps - 1; t >= 0; --t)
      if (tok_out[int64_t(t) * B + b] <= 0) k = t;
    cap_len[b] = k;
    atomicMax(&s_max, k);
  }
  __syncthreads();
  const int n = s_max;
  for (int b = threadIdx.x; b < B; b += blockDim.x) cap_len[b] = min(cap_len[b] + 2, n + 1);
  if (threadIdx.x == 0) n_out[0] = n;
}

// ------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------
static int check_dims(const coopcap_speaker* c) {
  CC_REQUIRE(c != nullptr, "speaker: null context");
  CC_REQUIRE(c->B > 0 && c->L > 0 && c->NL > 0, "speaker: empty batch B=%d L=%d NL=%d", c->B, c->L,
             c->NL);
  CC_REQUIRE(c->D % 8 == 0 && c->R % 8 == 0 && c->E % 8 == 0 && c->A % 8 == 0 && c->V1 % 4 == 0,
             "speaker: D,R,E,A must be multiples of 8 and V1 of 4");
  CC_REQUIRE(ATT_THREADS % (c->R / 8) == 0 && c->R / 8 <= ATT_THREADS,
             "speaker: R/8 must divide %d", ATT_THREADS);
  CC_REQUIRE(c->n_steps >= 0 && c->n_steps <= c->cap, "speaker: n_steps %d > cap %d", c->n_steps,
             c->cap);
  CC_REQUIRE(c->drop_p >= 0.f && c->drop_p < 1.f, "speaker: drop_p %f", c->drop_p);
  if (c->mode == COOPCAP_SAMPLE_PS_GUMBEL || c->mode == COOPCAP_SAMPLE_PS_MULTINOMIAL)
    CC_REQUIRE(c->soft16 && c->ps_sel && c->w_embed16 && c->V1 % 8 == 0,
               "speaker: partial-sampling modes need soft16, ps_sel, w_embed16 and V1 %% 8 == 0");
  CC_REQUIRE(c->ss_prob <= 0.f || (c->mode == COOPCAP_SAMPLE_MULTINOMIAL && c->forced),
             "speaker: scheduled sampling needs mode MULTINOMIAL and forced targets");
  return CC_OK;
}

size_t attention_smem_bytes(int A, int R, int L) {
  const int groups = ATT_THREADS / (R / 8);
  return sizeof(float) * (2 * A + 8 + ((L + 3) & ~3) + groups * R);
}

// ------------------------------------------------------------------------------------------
// one decode step's vocabulary layer: logits + sampling in the GEMM epilogue, then the per-row
// finish (logit_sample.cuh)
// ------------------------------------------------------------------------------------------
template <int MODE, bool INJ>
static int launch_logit_sample(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                               const LogitSampleParams& p, cudaStream_t s) {
  auto kern = logit_sample_kernel<MODE, INJ>;
  int rc;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), LsCfg::SMEM_BYTES))) return rc;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + LS_BN - 1) / LS_BN);
  const int grid = std::min(num_sms(), tiles);
  CC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(LS_THREADS), size_t(LsCfg::SMEM_BYTES), s, tmA, tmB, M,
                           N, K, p));
  return CC_OK;
}

static int logit_sample_step(const coopcap_speaker* c, int t, const bf16* out16_t, __half* z_t,
                             bf16* x_next, cudaStream_t s) {
  const int B = c->B, R = c->R, E = c->E, V1 = c->V1, XH = E + R;
  CC_REQUIRE(c->z16_all && c->ls_part && c->z_tgt, "speaker: z16_all / ls_part / z_tgt workspace missing");
  CC_REQUIRE(V1 % 8 == 0, "speaker: V1 must be a multiple of 8 (fp16 logit rows are written 16 bytes at a time)");
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, out16_t, 2, B, R, R, GEMM_BM, LsCfg::BK))) return rc;
  if ((rc = encode_tmap_2d(&tmB, c->w_logit16, 2, V1, R, R, LS_BN, LsCfg::BK))) return rc;
  LogitSampleParams p = {};
  p.bias = c->b_logit;
  p.noise = c->noise ? c->noise + int64_t(t) * B * V1 : nullptr;
  p.forced = c->forced ? c->forced + int64_t(t) * B : nullptr;
  p.ban = (c->no_repeat && t > 0) ? c->tok_out + int64_t(t - 1) * B : nullptr;    // AttModel.py:437-442
  p.z16 = z_t;
  p.part = c->ls_part;
  p.z_tgt = c->z_tgt;
  p.seed = c->seed;
  p.nstream = uint64_t(SITE_NOISE + t);
  p.inv_tau = c->inv_tau;
  p.store_pert = (c->store_perturbed && c->mode == COOPCAP_SAMPLE_ST_GUMBEL) ? 1 : 0;
  const bool inj = c->noise != nullptr;
  switch (c->mode) {
    case COOPCAP_SAMPLE_GREEDY:
      rc = launch_logit_sample<COOPCAP_SAMPLE_GREEDY, false>(tmA, tmB, B, V1, R, p, s); break;
    case COOPCAP_SAMPLE_NONE:
      rc = launch_logit_sample<COOPCAP_SAMPLE_NONE, false>(tmA, tmB, B, V1, R, p, s); break;
    case COOPCAP_SAMPLE_MULTINOMIAL:
      rc = inj ? launch_logit_sample<COOPCAP_SAMPLE_MULTINOMIAL, true>(tmA, tmB, B, V1, R, p, s)
               : launch_logit_sample<COOPCAP_SAMPLE_MULTINOMIAL, false>(tmA, tmB, B, V1, R, p, s);
      break;
    case COOPCAP_SAMPLE_ST_GUMBEL:
    case COOPCAP_SAMPLE_PS_GUMBEL:        // same epilogue: y = softmax((z + G) / tau), id = argmax
      rc = inj ? launch_logit_sample<COOPCAP_SAMPLE_ST_GUMBEL, true>(tmA, tmB, B, V1, R, p, s)
               : launch_logit_sample<COOPCAP_SAMPLE_ST_GUMBEL, false>(tmA, tmB, B, V1, R, p, s);
      break;
    case COOPCAP_SAMPLE_ST_MULTINOMIAL:
    case COOPCAP_SAMPLE_PS_MULTINOMIAL:   // same epilogue: softmax(z / tau) sums, id from the race
      rc = inj ? launch_logit_sample<COOPCAP_SAMPLE_ST_MULTINOMIAL, true>(tmA, tmB, B, V1, R, p, s)
               : launch_logit_sample<COOPCAP_SAMPLE_ST_MULTINOMIAL, false>(tmA, tmB, B, V1, R, p, s);
      break;
    default:
      set_last_error("speaker: unknown sampling mode %d", c->mode);
      return CC_ERR_ARG;
  }
  if (rc) return rc;
  // dense contraction; algorithmic bytes of the fused sampler: the fp16 logits written once
  prof_mark(PROF_LOGIT_SAMPLE, s, 2.0 * double(B) * V1 * R, 2.0 * (double(B) * R + double(V1) * R) + 2.0 * double(B) * V1);
  const int nrec = LS_RECS_PER_TILE * ((V1 + LS_BN - 1) / LS_BN);
  CC_CHECK_CUDA(launch_pdl(
      sample_finish_kernel, dim3(B), dim3(FIN_THREADS), 0, s, static_cast<const float*>(c->ls_part), nrec,
      static_cast<const float*>(c->z_tgt), c->mode, c->inv_tau, V1, p.forced,
      static_cast<const uint8_t*>(t > 0 ? c->unfinished + int64_t(t - 1) * B : nullptr),
      c->tok_raw + int64_t(t) * B, c->tok_out + int64_t(t) * B, c->tok_fed + int64_t(t + 1) * B,
      c->logp + int64_t(t) * B, c->lse + int64_t(t) * B, c->y_max + int64_t(t) * B,
      c->y_sum + int64_t(t) * B, c->unfinished + int64_t(t) * B, c->embed, E,
      static_cast<const uint8_t*>(c->keep_embed ? c->keep_embed + int64_t(t + 1) * B * E : nullptr),
      c->seed, uint64_t(SITE_DROP_EMBED + t + 1), c->drop_p, x_next, int64_t(XH), c->ss_prob,
      static_cast<const float*>(c->ss_u ? c->ss_u + int64_t(t) * B : nullptr), uint64_t(SITE_SCHED + t)));
  CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 32.0 * double(B) * nrec);
  return CC_OK;
}

// One attention step over the rows of `c` (c->B rows, regions c->att_off): s_t holds att_h at column
// 5R of every row.  Shared by the decode loop and the beam search (csrc/beam.cu), which calls it once
// per beam slot with that slot's rows.
int attention_fwd_launch(const coopcap_speaker* c, const float* s_t, bf16* att_res16_t, float* att_w_t,
                         cudaStream_t s, float* att_res32_t) {
  const int B = c->B, R = c->R, A = c->A, NS = 5 * R + A;
  int rc;
  if (A == 512 && R == 512) {
    if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_fwd4_kernel<512>), ATT4_SMEM))) return rc;
    CC_CHECK_CUDA(launch_pdl(attention_fwd4_kernel<512>, dim3(std::min(num_sms(), B)), dim3(ATT4_THREADS),
                             size_t(ATT4_SMEM), s, reinterpret_cast<const bf16*>(c->p_att16),
                             reinterpret_cast<const bf16*>(c->att_e16), c->att_off, c->L, c->att_order, s_t,
                             int64_t(NS), 5 * R, c->w_alpha, att_res16_t, att_w_t, B, att_res32_t));
  } else {
    const size_t att_smem = attention_smem_bytes(A, R, c->L);
    if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(attention_fwd_kernel), int(att_smem)))) return rc;
    attention_fwd_kernel<<<B, ATT_THREADS, att_smem, s>>>(
        reinterpret_cast<const bf16*>(c->p_att16), reinterpret_cast<const bf16*>(c->att_e16), c->att_off, c->L,
        s_t, NS, 5 * R, c->w_alpha, att_res16_t, att_w_t, A, R);
  }
  // algorithmic bytes: p_att + att_e read once (bf16), att_h in, att_res + weights out
  CC_LAUNCH_CHECK_K(PROF_ATT_FWD, s, 0.0, 2.0 * c->NL * (A + R) + 4.0 * B * A + 2.0 * B * R + 4.0 * c->NL);
  return CC_OK;
}

// maxout-LSTM pointwise step over `rows` rows (AttModel.py:515-531)
int lstm_fwd_launch(const coopcap_speaker* c, const float* s_t, const float* u_t, const float* c_prev,
                    float* c_next, bf16* h_out16, int64_t ld_h, bf16* out16_t, const uint8_t* keep,
                    uint64_t site, float drop_p, int rows, cudaStream_t s) {
  const int R = c->R, NS = 5 * R + c->A;
  const int n = rows * (R / 4);
  CC_CHECK_CUDA(launch_pdl(lstm_fwd_kernel, dim3((n + 255) / 256), dim3(256), 0, s, s_t, int64_t(NS), u_t, c_prev,
                           c_next, h_out16, ld_h, out16_t, keep, c->seed, site, drop_p, rows, R));
  CC_LAUNCH_CHECK_K(PROF_LSTM, s, 0.0, 0.0);
  return CC_OK;
}

int speaker_prologue_fwd(const coopcap_speaker* c, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  if (!c->att_prepacked) {
    CC_REQUIRE(c->att_feats != nullptr, "speaker: att_feats is null and att16 is not pre-packed");
    CC_CHECK_CUDA(launch_pdl(pack_att_kernel<1>, dim3(c->NL), dim3(PACK_THREADS), size_t(0), s, c->att_feats, c->att_off, c->B, c->L, c->D, c->NL,
                                          reinterpret_cast<bf16*>(c->att16)));
    CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 6.0 * c->NL * c->D);
  }
  EpiStoreParams ep = {};
  ep.alpha = 1.f;
  ep.bias = c->b_att_embed;
  ep.relu = 1;
  ep.C16 = reinterpret_cast<bf16*>(c->att_e16);
  ep.ldc16 = c->R;
  if (c->drop_p > 0.f) {
    ep.drop_p = c->drop_p;
    if (c->keep_att) { ep.keep = c->keep_att; ep.ld_keep = c->R; }
    else { ep.philox_dropout = 1; ep.seed = c->seed; ep.stream = SITE_DROP_ATT; }
  }
  rc = gemm_run(0, 0, 0, c->att16, c->D, c->w_att_embed16, c->D, c->NL, c->R, c->D, 1, 0, ep, s);
  if (rc) return rc;
  EpiStoreParams e2 = {};
  e2.alpha = 1.f;
  e2.bias = c->b_ctx2att;
  e2.C16 = reinterpret_cast<bf16*>(c->p_att16);
  e2.ldc16 = c->A;
  return gemm_run(0, 0, 0, c->att_e16, c->R, c->w_ctx2att16, c->R, c->NL, c->A, c->R, 1, 0, e2, s);
}

int speaker_decode_fwd(const coopcap_speaker* c, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  const int B = c->B, R = c->R, E = c->E, A = c->A, V1 = c->V1;
  const int NS = 5 * R + A, XH = E + R;
  bf16* xh16 = reinterpret_cast<bf16*>(c->xh16);
  bf16* att_res16 = reinterpret_cast<bf16*>(c->att_res16);
  bf16* out16 = reinterpret_cast<bf16*>(c->out16);
  CC_CHECK_CUDA(launch_pdl(start_step_kernel, dim3(B), dim3(128), size_t(0), s, c->embed, c->start_token, c->start_tokens, B, E, R, c->keep_embed, c->seed,
                                      c->drop_p, xh16, c->c_all, c->tok_fed));
  CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 0.0);
  for (int t = 0; t < c->n_steps; ++t) {
    float* s_t = c->s_all + int64_t(t) * B * NS;
    float* u_t = c->u_all + int64_t(t) * B * 2 * R;
    // gates + att_h:  [x_t | h_{t-1}] . w_cat^T + b_cat
    EpiStoreParams e1 = {};
    e1.alpha = 1.f; e1.bias = c->b_cat; e1.C = s_t; e1.ldc = NS;
    rc = gemm_run(0, 0, 0, xh16 + int64_t(t) * B * XH, XH, c->w_cat16, XH, B, NS, XH, 1, 0, e1, s);
    if (rc) return rc;
    if ((rc = attention_fwd_launch(c, s_t, att_res16 + int64_t(t) * B * R, c->att_w + int64_t(t) * c->NL, s,
                                   c->att_res32 ? c->att_res32 + int64_t(t) * B * R : nullptr)))
      return rc;
    if (cell_step_ok<LstmCell>(R, R)) {
      // a2c GEMM with the maxout-LSTM update as its epilogue (csrc/cell_step.cuh)
      LstmStepParams lp = {};
      lp.b_a2c = c->b_a2c; lp.s = s_t; lp.lds = NS; lp.c_prev = c->c_all + int64_t(t) * B * R; lp.u = u_t;
      lp.c_next = c->c_all + int64_t(t + 1) * B * R; lp.h_dst = xh16 + int64_t(t + 1) * B * XH + E; lp.ld_h = XH;
      lp.out16 = out16 + int64_t(t) * B * R;
      lp.keep = c->keep_core ? c->keep_core + int64_t(t) * B * R : nullptr;
      lp.seed = c->seed; lp.stream = uint64_t(SITE_DROP_CORE + t); lp.drop_p = c->drop_p;
      if ((rc = launch_cell_step<LstmCell>(att_res16 + int64_t(t) * B * R, R, c->w_a2c16, B, R, R, lp, s))) return rc;
    } else {
      EpiStoreParams e2 = {};
      e2.alpha = 1.f; e2.bias = c->b_a2c; e2.C = u_t; e2.ldc = 2 * R;
      rc = gemm_run(0, 0, 0, att_res16 + int64_t(t) * B * R, R, c->w_a2c16, R, B, 2 * R, R, 1, 0, e2, s);
      if (rc) return rc;
      if ((rc = lstm_fwd_launch(c, s_t, u_t, c->c_all + int64_t(t) * B * R, c->c_all + int64_t(t + 1) * B * R,
                                xh16 + int64_t(t + 1) * B * XH + E, int64_t(XH), out16 + int64_t(t) * B * R,
                                c->keep_core ? c->keep_core + int64_t(t) * B * R : nullptr,
                                uint64_t(SITE_DROP_CORE + t), c->drop_p, B, s)))
        return rc;
    }
    __half* z_t = reinterpret_cast<__half*>(c->z16_all) + int64_t(t) * B * V1;
    const bool ps = (c->mode == COOPCAP_SAMPLE_PS_GUMBEL || c->mode == COOPCAP_SAMPLE_PS_MULTINOMIAL);
    bf16* x_next = (t + 1 < c->n_steps) ? xh16 + int64_t(t + 1) * B * XH : nullptr;
    if ((rc = logit_sample_step(c, t, out16 + int64_t(t) * B * R, z_t, ps ? nullptr : x_next, s))) return rc;
    if (ps) {
      bf16* v_t = reinterpret_cast<bf16*>(c->soft16) + int64_t(t) * B * V1;
      CC_CHECK_CUDA(launch_pdl(
          ps_vec_kernel, dim3(B), dim3(256), 0, s, z_t, V1, c->mode, c->inv_tau,
          c->noise ? c->noise + int64_t(t) * B * V1 : nullptr, c->seed, uint64_t(SITE_NOISE + t),
          c->y_max + int64_t(t) * B, c->y_sum + int64_t(t) * B, c->tok_fed + int64_t(t + 1) * B,
          c->ps_prob, c->part_u ? c->part_u + int64_t(t) * B : nullptr, uint64_t(SITE_PARTIAL + t),
          v_t, c->ps_sel + int64_t(t) * B));
      CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 6.0 * B * V1);
      if (x_next) {
        // x_{t+1} = dropout(relu(v_t . embed))                        (AttModel.py:395-397)
        EpiStoreParams e = {};
        e.alpha = 1.f; e.relu = 1; e.C16 = x_next; e.ldc16 = XH;
        if (c->drop_p > 0.f) {
          e.drop_p = c->drop_p;
          if (c->keep_embed) { e.keep = c->keep_embed + int64_t(t + 1) * B * E; e.ld_keep = E; }
          else { e.philox_dropout = 1; e.seed = c->seed; e.stream = SITE_DROP_EMBED + t + 1; }
        }
        rc = gemm_run(0, 0, 1, v_t, V1, c->w_embed16, E, B, E, V1, 1, 64, e, s);
        if (rc) return rc;
      }
      ps_mask_kernel<<<B, 256, 0, s>>>(c->unfinished + int64_t(t) * B, V1, v_t);
      CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 0.0);
    }
  }
  if (c->n_out && c->cap_len) {
    CC_CHECK_CUDA(launch_pdl(caption_summary_kernel, dim3(1), dim3(1024), size_t(0), s, c->tok_out, B, c->n_steps, c->n_out, c->cap_len));
    CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  }
  return CC_OK;
}

}  // namespace coopcap

extern "C" {

int coopcap_set_cast_multi(int on) {
  coopcap::set_cast_multi(on);
  return coopcap::CC_OK;
}

int coopcap_speaker_pack_weights(const coopcap_speaker_pack* p, coopcap_stream_t stream) {
  using namespace coopcap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CC_REQUIRE(p != nullptr, "speaker_pack: null");
  const int R = p->R, E = p->E, A = p->A, D = p->D, XH = E + R;
  bf16* wc = reinterpret_cast<bf16*>(p->w_cat16);
  int rc;
  CC_CHECK_CUDA(cudaMemset2DAsync(wc + int64_t(5 * R) * XH, XH * sizeof(bf16), 0, E * sizeof(bf16),
                                  A, s));
  {
    const float* src[7] = {p->w_att_embed, p->w_ctx2att, p->w_i2h, p->w_h2h, p->w_h2att, p->w_a2c, p->w_logit};
    const int64_t rows[7] = {R, A, 5 * R, 5 * R, A, 2 * R, p->V1};
    const int cols[7] = {D, R, E, R, R, R, R};
    void* dst[7] = {p->w_att_embed16, p->w_ctx2att16, wc, wc + E, wc + int64_t(5 * R) * XH + E,
                    p->w_a2c16, p->w_logit16};
    const int64_t ld[7] = {D, R, XH, XH, XH, R, R};
    if ((rc = cast_blocks(7, src, rows, cols, dst, ld, s))) return rc;
  }
  const int n = 5 * R + A;
  bias_cat_kernel<<<(n + 255) / 256, 256, 0, s>>>(p->b_i2h, p->b_h2h, p->b_h2att, 5 * R, A, p->b_cat);
  CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 0.0);
  return CC_OK;
}

int coopcap_pack_att_from_host(const float* att_feats_pinned, const int* att_off, int B, int L, int D,
                               int NL, void* att16, int ctas, coopcap_stream_t stream) {
  using namespace coopcap;
  CC_REQUIRE(att_feats_pinned && att16 && B > 0 && L > 0 && NL > 0 && D % 4 == 0,
             "pack_att_from_host: bad arguments");
  void* dptr = nullptr;
  // the host buffer must be pinned + mapped (UVA): ask the runtime for its device alias
  CC_CHECK_CUDA(cudaHostGetDevicePointer(&dptr, const_cast<float*>(att_feats_pinned), 0));
  if (ctas <= 0) ctas = 64;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // rows in flight per CTA (tuning knob COOPCAP_PACK_ROWS = 1, 2 or 4).  Measured inside the
  // end-to-end step (upload overlapping compute), 64 CTAs: 1 row 10.5 ms/step, 2 rows 12.3, 4 rows
  // 12.8; 128 CTAs are worse at every setting -- more requests in flight do not help this path.
  static const int rows_per_item = [] {
    const char* e = getenv("COOPCAP_PACK_ROWS");
    const int v = e ? atoi(e) : 1;
    return (v == 1 || v == 2 || v == 4) ? v : 1;
  }();
  const int items = (NL + rows_per_item - 1) / rows_per_item;
  if (ctas > items) ctas = items;
  const float* src = reinterpret_cast<const float*>(dptr);
  bf16* dst = reinterpret_cast<bf16*>(att16);
  if (rows_per_item == 1) CC_CHECK_CUDA(launch_pdl(pack_att_kernel<1>, dim3(ctas), dim3(PACK_THREADS), size_t(0), s, src, att_off, B, L, D, NL, dst));
  else if (rows_per_item == 2) pack_att_kernel<2><<<ctas, PACK_THREADS, 0, s>>>(src, att_off, B, L, D, NL, dst);
  else pack_att_kernel<4><<<ctas, PACK_THREADS, 0, s>>>(src, att_off, B, L, D, NL, dst);
  CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 4.0 * NL * D);
  return CC_OK;
}

int coopcap_speaker_prologue_fwd(const coopcap_speaker* ctx, coopcap_stream_t stream) {
  return coopcap::speaker_prologue_fwd(ctx, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_speaker_decode_fwd(const coopcap_speaker* ctx, coopcap_stream_t stream) {
  return coopcap::speaker_decode_fwd(ctx, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
