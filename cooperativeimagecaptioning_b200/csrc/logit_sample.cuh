// Vocabulary logits + sampling in ONE kernel (models/AttModel.py:444 `logit` followed by
// models/gumbel.py:6-30, models/multinomial.py:4-27, AttModel.py:328-365,403-409).
//
//   z[b, v] = out[b, :] . W_logit[v, :] + bias[v]                  (tcgen05, fp32 in TMEM)
//
// The epilogue never writes fp32 logits.  Each epilogue thread owns one batch row and 64 columns of
// the 128 x 256 accumulator tile and, straight out of TMEM, folds them into a partial record:
//   (m1, s1)  running max / sum of exp for log-sum-exp(z)                       -> lse, log-probs
//   (m2, s2)  the same for the relaxed sample's scores (z + G) / tau or z / tau -> y_max, y_sum
//   (bv, bi, bz)  best perturbed score, its column and its raw logit            -> sampled id
// with the Gumbel / exponential-race noise generated in registers (Philox keyed by the element's
// position, so the backward pass regenerates the same noise) or read from an injected tensor
// (parity tests).  The logits leave the SM once, as fp16, and every quantity
// above is computed from those rounded values, so the backward pass -- which re-reads them to
// rebuild softmax(z) and y -- differentiates exactly the function that ran.
// `sample_finish_kernel` merges the 4 x 38 partials of a row, applies the reference's token
// bookkeeping and writes the next step's input embedding.
//
// Against the separate logit GEMM + row-streaming sampler this replaces (r1): per decode step one
// 38.9 MB fp32 store, one 38.9 MB read and a launch disappear; backward reads half the bytes.
#pragma once
#include <cuda_fp16.h>
#include "gemm.cuh"
#include "speaker_kernels.cuh"

namespace coopcap {

constexpr int LS_BN = 256;
constexpr int LS_REC = 8;                 // floats per partial record
constexpr int LS_EPI_WARPS = 16;          // four per TMEM lane quarter: 64 columns of the tile each
constexpr int LS_THREADS = (2 + LS_EPI_WARPS) * 32;
constexpr int LS_RECS_PER_TILE = LS_EPI_WARPS / 4;   // records a row gets from one column tile
constexpr int LS_SUB = 16;                // columns folded per register batch

struct LsCfg {                            // bf16 operands, both K-major, BN = 256
  static constexpr int EB = 2, BK = 64, UK = 16, MN_ATOM = 64;
  static constexpr int A_BYTES = GEMM_BM * 128;
  static constexpr int B_BYTES = LS_BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2 * LS_BN * 4;                   // double-buffered bias slice
  static constexpr int STAGES_RAW =
      (GEMM_SMEM_TOTAL - GEMM_SMEM_EXTRA - 1024 - BIAS_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_SMEM_EXTRA + BIAS_BYTES + 1024;
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

struct LogitSampleParams {
  const float* bias;        // [N]
  const float* noise;       // injected noise [M, N] (uniforms for Gumbel, Exp(1) draws for the race) or null
  const int64_t* forced;    // [M] ids whose raw logit is wanted (teacher forcing / replay) or null
  const int64_t* ban;       // [M] id whose logit becomes -inf (decoding_constraint) or null
  __half* z16;              // [M, N] the logits, fp16
  float* part;              // [M][LS_RECS_PER_TILE * num_n][LS_REC] partial records
  float* z_tgt;             // [M] logit of forced[b] (written only when forced != null)
  uint64_t seed, nstream;
  float inv_tau;
  int store_pert;           // ST-Gumbel only: z16 receives the PERTURBED logits z + G (see coopcap_speaker)
};

__device__ __forceinline__ void epi_barrier() {       // the epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(LS_EPI_WARPS * 32) : "memory");
}

// (max, sum-of-exp) in the base-2 domain: m = max_j(scale * x_j), s = sum_j 2^(scale * x_j - m)
struct Lse2 {
  float m, s;
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; }
  // fold LS_SUB values x_j * scale (scale > 0)
  __device__ __forceinline__ void fold(const float (&x)[LS_SUB], float scale) {
    float a = fmaxf(x[0], x[1]), b = fmaxf(x[2], x[3]), c = fmaxf(x[4], x[5]), d = fmaxf(x[6], x[7]);
#pragma unroll
    for (int j = 8; j < LS_SUB; j += 4) {
      a = fmaxf(a, x[j]); b = fmaxf(b, x[j + 1]); c = fmaxf(c, x[j + 2]); d = fmaxf(d, x[j + 3]);
    }
    const float nm = fmaxf(fmaxf(fmaxf(a, b), fmaxf(c, d)) * scale, m);
    if (nm == -INFINITY) return;                       // nothing but -inf so far
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;      // four chains: the adds are latency-bound otherwise
#pragma unroll
    for (int j = 0; j < LS_SUB; j += 4) {
      t0 += ex2_ftz(fmaf(x[j], scale, -nm)); t1 += ex2_ftz(fmaf(x[j + 1], scale, -nm));
      t2 += ex2_ftz(fmaf(x[j + 2], scale, -nm)); t3 += ex2_ftz(fmaf(x[j + 3], scale, -nm));
    }
    s = fmaf(s, ex2_ftz(m - nm), (t0 + t1) + (t2 + t3));
    m = nm;
  }
  __device__ __forceinline__ void merge(float m2, float s2) {
    const float nm = fmaxf(m, m2);
    if (nm == -INFINITY) return;
    s = s * ex2_ftz(m - nm) + s2 * ex2_ftz(m2 - nm);
    m = nm;
  }
};

template <int MODE, bool INJ>
__global__ void __launch_bounds__(LS_THREADS, 1)
logit_sample_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    int M, int N, int K, LogitSampleParams p) {
  using Cfg = LsCfg;
  constexpr int BN = LS_BN;
  constexpr int STAGES = Cfg::STAGES;
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr bool gum = (MODE == COOPCAP_SAMPLE_ST_GUMBEL || MODE == COOPCAP_SAMPLE_PS_GUMBEL);
  constexpr bool race = (MODE == COOPCAP_SAMPLE_MULTINOMIAL || MODE == COOPCAP_SAMPLE_ST_MULTINOMIAL ||
                         MODE == COOPCAP_SAMPLE_PS_MULTINOMIAL);
  constexpr bool st = gum || MODE == COOPCAP_SAMPLE_ST_MULTINOMIAL || MODE == COOPCAP_SAMPLE_PS_MULTINOMIAL;
  constexpr bool pick = (MODE != COOPCAP_SAMPLE_NONE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* s_bias = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + GEMM_SMEM_EXTRA);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int nkb = (K + Cfg::BK - 1) / Cfg::BK;
  const int num_tiles = num_m * num_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], LS_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0)
      gemm_producer_role<Cfg, BN, 0, 0>(&tmA, &tmB, sA, sB, full_bar, empty_bar, num_m, num_n,
                                        num_tiles, nkb, nkb);
  } else if (warp == 1) {
    if (lane == 0)
      gemm_mma_role<0, Cfg, BN, 0, 0>(sA, sB, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base,
                                      num_m, num_n, num_tiles, nkb, nkb);
  } else {
    // ============ epilogue: 16 warps, four per TMEM lane quarter, 64 columns of the tile each ============
    // (the sampler's instruction stream is long dependent chains -- Philox rounds, MUFU -> FMA ->
    // MUFU: with two warps per scheduler it issued 28 % of the cycles, ncu r2f; four hide it)
    const int q = warp & 3;              // TMEM lane quarter (rows q*32 .. q*32+31 of the tile)
    const int g = (warp - 2) >> 2;       // this warp's 64-column group of the tile
    const int et = (warp - 2) * 32 + lane;
    const float k_tau2 = p.inv_tau * LOG2E;            // logits -> base-2 score units
    const int nv4 = N >> 2;
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % num_m, n_blk = t / num_m;
      const int row0 = m_blk * GEMM_BM + q * 32;
      const int row = row0 + lane;
      const int n0 = n_blk * BN;
      float* sb = s_bias + as * BN;
      if (et < BN) sb[et] = (n0 + et < N) ? __ldg(p.bias + n0 + et) : 0.f;
      const bool live = row < M;
      const int tgt = (p.forced && live) ? int(p.forced[row]) : -1;
      const int ban = (p.ban && live) ? int(p.ban[row]) : -1;
      const float* nrow = INJ ? p.noise + int64_t(live ? row : 0) * N : nullptr;
      __half* zrow = p.z16 + int64_t(live ? row : 0) * N;
      Lse2 l1, l2;
      l1.init();
      l2.init();
      float bv = -INFINITY, bz = 0.f, zt = 0.f;
      int bi = 0x7fffffff;
      bool have_zt = false;
      epi_barrier();                                   // bias slice staged
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + as * BN + g * 64;
      const bool group_live = (n0 + g * 64 < N) && (row0 < M);       // warp-uniform
#pragma unroll 1
      for (int c = 0; c < 64 / LS_SUB; ++c) {
        const int tc = g * 64 + c * LS_SUB;            // first column within the tile
        const int col0 = n0 + tc;
        float x[LS_SUB];
        tmem_ld16(t_addr + c * LS_SUB, x);
        tmem_ld_wait();
        if (c == 64 / LS_SUB - 1) {                    // all of this thread's columns are out of TMEM
          tc_fence_before();
          mbar_arrive(&tempty_bar[as]);
        }
        const int ncols = N - col0;                    // >= LS_SUB: full batch; <= 0: nothing valid
        if (!group_live || ncols <= 0) continue;       // warp-uniform
        {
          const float4* b4 = reinterpret_cast<const float4*>(sb + tc);
#pragma unroll
          for (int j4 = 0; j4 < LS_SUB / 4; ++j4) {
            const float4 bb = b4[j4];
            x[4 * j4] += bb.x; x[4 * j4 + 1] += bb.y; x[4 * j4 + 2] += bb.z; x[4 * j4 + 3] += bb.w;
          }
        }
        if (ncols < LS_SUB) {
#pragma unroll
          for (int j = 0; j < LS_SUB; ++j)
            if (j >= ncols) x[j] = -INFINITY;
        }
        if (__any_sync(0xffffffffu, unsigned(ban - col0) < unsigned(LS_SUB))) {
          const int d = ban - col0;
#pragma unroll
          for (int j = 0; j < LS_SUB; ++j)
            if (j == d) x[j] = -INFINITY;
        }
        // ---- the logits leave as fp16, straight from registers: 32 contiguous bytes per thread.
        // From here on the layer's logits ARE these rounded values (11-bit mantissa, finer than the
        // bf16 operands that produced them): log-sum-exp, log-probs, the relaxed sample y and the
        // drawn id are all functions of exactly what backward re-reads, so softmax(z) and y sum to
        // one there and dz is the gradient of the function that ran.
#pragma unroll
        for (int jj = 0; jj < LS_SUB / 8; ++jj) {
          __half2 h[4];
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            h[k2] = __floats2half2_rn(x[8 * jj + 2 * k2], x[8 * jj + 2 * k2 + 1]);
            const float2 f = __half22float2(h[k2]);
            x[8 * jj + 2 * k2] = f.x;
            x[8 * jj + 2 * k2 + 1] = f.y;
          }
          uint4 o;
          o.x = *reinterpret_cast<const uint32_t*>(&h[0]);
          o.y = *reinterpret_cast<const uint32_t*>(&h[1]);
          o.z = *reinterpret_cast<const uint32_t*>(&h[2]);
          o.w = *reinterpret_cast<const uint32_t*>(&h[3]);
          if (live && 8 * jj < ncols && !(gum && p.store_pert))
            *reinterpret_cast<uint4*>(zrow + col0 + 8 * jj) = o;   // N % 8 == 0
        }
        if (__any_sync(0xffffffffu, unsigned(tgt - col0) < unsigned(LS_SUB))) {
          const int d = tgt - col0;
#pragma unroll
          for (int j = 0; j < LS_SUB; ++j)
            if (j == d) { zt = x[j]; have_zt = true; }
        }
        // ---- noise
        float a2[LS_SUB];                              // perturbed scores (arg-max key), base-2 units
        if constexpr (gum || race) {
#pragma unroll
          for (int j4 = 0; j4 < LS_SUB / 4; ++j4) {
            float u[4];
            if constexpr (INJ) {
              // (the row's tail past N belongs to the next row -- or to nobody on the last one)
              const float4 tq = (4 * j4 < ncols) ? *reinterpret_cast<const float4*>(nrow + col0 + 4 * j4)
                                                 : make_float4(0.5f, 0.5f, 0.5f, 0.5f);
              u[0] = tq.x; u[1] = tq.y; u[2] = tq.z; u[3] = tq.w;
            } else {
              const uint4 r = Philox::gen_r<NOISE_ROUNDS>(
                  p.seed, p.nstream, uint64_t(row) * nv4 + uint64_t((col0 >> 2) + j4));
              u[0] = Philox::u01(r.x); u[1] = Philox::u01(r.y);
              u[2] = Philox::u01(r.z); u[3] = Philox::u01(r.w);
            }
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int j = 4 * j4 + qq;
              if constexpr (INJ) {
                // libm-accurate transforms: parity with the reference's torch.log on the same draws
                if constexpr (gum) a2[j] = (x[j] + gumbel_of(u[qq], false)) * k_tau2;
                else a2[j] = (x[j] * p.inv_tau - logf(u[qq])) * LOG2E;
              } else {
                // -log2(e), e ~ Exp(1): Gumbel noise is -ln(e), the race adds -ln(e) as well
                const float nl = lg2_ftz(neg_log_fast(u[qq]));
                if constexpr (gum) a2[j] = fmaf(-p.inv_tau, nl, x[j] * k_tau2);
                else a2[j] = fmaf(x[j], k_tau2, -nl);
              }
            }
          }
        }
        if constexpr (MODE == COOPCAP_SAMPLE_GREEDY) {
#pragma unroll
          for (int j = 0; j < LS_SUB; ++j) a2[j] = x[j];
        }
        // ---- arg-max of the perturbed score (first index wins ties, like torch.max)
        if constexpr (pick) {
#pragma unroll
          for (int j = 0; j < LS_SUB; ++j)
            if (a2[j] > bv) { bv = a2[j]; bi = col0 + j; bz = x[j]; }
        }
        if constexpr (gum) {
          if (p.store_pert) {
            // what leaves the SM is z + G, rounded to fp16, and the relaxed sample y is computed from
            // exactly those rounded values: backward rebuilds y from them without regenerating a
            // single noise value (the id above was drawn from the unrounded scores)
            const float inv_k = 1.f / k_tau2;
#pragma unroll
            for (int jj = 0; jj < LS_SUB / 8; ++jj) {
              __half2 h[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                h[k2] = __floats2half2_rn(a2[8 * jj + 2 * k2] * inv_k, a2[8 * jj + 2 * k2 + 1] * inv_k);
                const float2 f = __half22float2(h[k2]);
                a2[8 * jj + 2 * k2] = f.x * k_tau2;
                a2[8 * jj + 2 * k2 + 1] = f.y * k_tau2;
              }
              uint4 o;
              o.x = *reinterpret_cast<const uint32_t*>(&h[0]);
              o.y = *reinterpret_cast<const uint32_t*>(&h[1]);
              o.z = *reinterpret_cast<const uint32_t*>(&h[2]);
              o.w = *reinterpret_cast<const uint32_t*>(&h[3]);
              if (live && 8 * jj < ncols) *reinterpret_cast<uint4*>(zrow + col0 + 8 * jj) = o;
            }
          }
        }
        // ---- log-sum-exp of the logits, and of the relaxed sample's scores
        l1.fold(x, LOG2E);
        if constexpr (gum) l2.fold(a2, 1.f);
        else if constexpr (st) l2.fold(x, k_tau2);
      }
      if (live) {
        float4* rec = reinterpret_cast<float4*>(
            p.part + (int64_t(row) * (LS_RECS_PER_TILE * num_n) + (n_blk * LS_RECS_PER_TILE + g)) * LS_REC);
        rec[0] = make_float4(l1.m, l1.s, l2.m, l2.s);
        rec[1] = make_float4(bv, __int_as_float(bi), bz, 0.f);
        if (have_zt) p.z_tgt[row] = zt;
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Merge the partials of one row, then the reference's per-step bookkeeping:
//   it = sampled / forced id; unfinished &= it > 0; seq = it * unfinished    (AttModel.py:403-409)
//   sampleLogprobs = logprobs.gather(it)                                     (:341,:357,:365)
// and the next step's input x = dropout(relu(embed[it]))                     (:74-76,:326-327)
// One CTA of 128 threads per row.
// ------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 128;

__global__ void __launch_bounds__(FIN_THREADS)
sample_finish_kernel(const float* __restrict__ part, int nrec, const float* __restrict__ z_tgt,
                     int mode, float inv_tau, int V1, const int64_t* __restrict__ forced,
                     const uint8_t* __restrict__ unf_prev, int64_t* __restrict__ tok_raw,
                     int64_t* __restrict__ tok_out, int64_t* __restrict__ tok_fed_next,
                     float* __restrict__ logp, float* __restrict__ lse_o, float* __restrict__ ymax_o,
                     float* __restrict__ ysum_o, uint8_t* __restrict__ unf,
                     // next-step input
                     const float* __restrict__ embed, int E, const uint8_t* __restrict__ keep_embed_next,
                     uint64_t seed, uint64_t estream, float drop_p, __nv_bfloat16* __restrict__ xh_next,
                     int64_t ld_xh,
                     // scheduled sampling (AttModel.py:119-131)
                     float ss_prob, const float* __restrict__ ss_u, uint64_t ss_stream) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_rec[FIN_THREADS / 32][LS_REC];
  __shared__ int64_t s_fed;
  constexpr float LN2 = 0.6931471805599453f;
  const int b = blockIdx.x;
  const float* pr = part + int64_t(b) * nrec * LS_REC;
  Lse2 l1, l2;
  l1.init();
  l2.init();
  float bv = -INFINITY, bz = 0.f;
  int bi = 0x7fffffff;
  for (int r = threadIdx.x; r < nrec; r += FIN_THREADS) {
    const float4 a = *reinterpret_cast<const float4*>(pr + r * LS_REC);
    const float4 c = *reinterpret_cast<const float4*>(pr + r * LS_REC + 4);
    l1.merge(a.x, a.y);
    l2.merge(a.z, a.w);
    const int oi = __float_as_int(c.y);
    if (c.x > bv || (c.x == bv && oi < bi)) { bv = c.x; bi = oi; bz = c.z; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l1.merge(__shfl_xor_sync(0xffffffffu, l1.m, o), __shfl_xor_sync(0xffffffffu, l1.s, o));
    l2.merge(__shfl_xor_sync(0xffffffffu, l2.m, o), __shfl_xor_sync(0xffffffffu, l2.s, o));
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o), oz = __shfl_xor_sync(0xffffffffu, bz, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bz = oz; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_rec[warp][0] = l1.m; s_rec[warp][1] = l1.s; s_rec[warp][2] = l2.m; s_rec[warp][3] = l2.s;
    s_rec[warp][4] = bv; s_rec[warp][5] = __int_as_float(bi); s_rec[warp][6] = bz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < FIN_THREADS / 32; ++w) {
      l1.merge(s_rec[w][0], s_rec[w][1]);
      l2.merge(s_rec[w][2], s_rec[w][3]);
      const float ov = s_rec[w][4];
      const int oi = __float_as_int(s_rec[w][5]);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bz = s_rec[w][6]; }
    }
    const float lse = (l1.m + log2f(l1.s)) * LN2;
    // a row of NaN logits (diverged training) wins no comparison: emit EOS, its NaN log-prob tells
    // the host; ids handed in by the caller are range-checked before they index anything
    if (unsigned(bi) >= unsigned(V1)) { bi = 0; bz = __int_as_float(0x7fc00000); }
    const int64_t raw = (mode == COOPCAP_SAMPLE_NONE) ? 0 : int64_t(bi);
    int64_t tgt = forced ? forced[b] : raw;
    const bool tgt_ok = uint64_t(tgt) < uint64_t(V1);
    if (!tgt_ok) tgt = 0;
    int64_t fed = tgt;
    if (ss_prob > 0.f) {               // scheduled sampling: feed the drawn id instead of the target
      const float u = ss_u ? ss_u[b] : Philox::u01(Philox::gen(seed, ss_stream, uint64_t(b)).x);
      if (u < ss_prob) fed = raw;
    }
    const bool up = unf_prev ? (unf_prev[b] != 0) : true;
    const bool un = up && (fed > 0);                 // AttModel.py:403-406
    tok_raw[b] = raw;
    tok_out[b] = un ? fed : 0;                       // :409
    tok_fed_next[b] = fed;
    const float z_of_tgt = forced ? z_tgt[b] : bz;
    logp[b] = tgt_ok ? z_of_tgt - lse : __int_as_float(0x7fc00000);
    lse_o[b] = lse;
    if (mode == COOPCAP_SAMPLE_PS_MULTINOMIAL) {
      // y = exp(log_softmax(z) / tau), unnormalised for tau != 1  (multinomial_soft.py:12-15)
      ymax_o[b] = lse * inv_tau;
      ysum_o[b] = 1.f;
    } else {
      ymax_o[b] = l2.m * LN2;                        // back to natural-log units (st_bwd_kernel)
      ysum_o[b] = l2.s;
    }
    unf[b] = un ? 1 : 0;
    s_fed = fed;
  }
  __syncthreads();
  if (xh_next) {
    embed_row(embed, s_fed, E, keep_embed_next ? keep_embed_next + int64_t(b) * E : nullptr, seed,
              estream, int64_t(b) * E, drop_p, xh_next + int64_t(b) * ld_xh);
  }
}

}  // namespace coopcap
