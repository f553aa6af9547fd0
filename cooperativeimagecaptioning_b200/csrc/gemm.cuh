// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = sum_k A[m,k] * B[n,k]        (fp32 accumulation in TMEM)
//
// * operands: bf16 (kind::f16) or fp32 read as tf32 (kind::tf32); each operand either K-major
//   (row-major [rows, K]) or MN-major (row-major [K, rows]) so that forward (X W^T), dgrad (dY W)
//   and wgrad (dY^T X) all run on the same kernel without transposed copies.
// * TMA (128-byte swizzle) -> smem ring -> one elected thread issues tcgen05.mma -> fp32
//   accumulators double-buffered in TMEM -> 4 epilogue warps read them back with tcgen05.ld
//   (one thread per output row) and run a fused epilogue functor.
// * warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue.
#pragma once
#include "common.cuh"

namespace coopcap {

constexpr int GEMM_BM = 128;
constexpr int GEMM_THREADS = 320;              // TMA warp + MMA warp + 8 epilogue warps
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_SMEM_TOTAL = 232448;         // 227 KB opt-in limit per CTA
constexpr int GEMM_SMEM_EXTRA = 2048;           // barriers + alignment slack
// epilogue staging: per epilogue warp EPI_BUFS buffers of [32 rows][128 B] fp32, 128B-swizzled
// (chunk j of row r lives at chunk j ^ (r & 7)); read by TMA stores or by the manual store path

template <int KIND, int BN>
struct GemmCfg {
  static constexpr int EB = (KIND == 0) ? 2 : 4;   // element bytes
  static constexpr int BK = 128 / EB;              // elements per 128-byte swizzle row
  static constexpr int UK = 32 / EB;               // K per tcgen05.mma
  static constexpr int MN_ATOM = 128 / EB;         // elements along MN per MN-major atom
  static constexpr int A_BYTES = GEMM_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BUFS = (BN <= 192) ? 2 : 1;
  static constexpr int EPI_STAGE_BYTES = GEMM_EPI_WARPS * EPI_BUFS * 4096;
  static constexpr int STAGES_RAW =
      (GEMM_SMEM_TOTAL - GEMM_SMEM_EXTRA - 1024 - EPI_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_SMEM_EXTRA + EPI_STAGE_BYTES + 1024;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32,256]");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

// ------------------------------------------------------------------------------------------
// Generic store epilogue: v = alpha*acc (+bias[col]) (+relu) ; row-major fp32 and/or bf16 output,
// optional transposed bf16 output; store / read-add-write / atomicAdd.
// ------------------------------------------------------------------------------------------
struct EpiStoreParams {
  float* C;              // [M, ldc] fp32 or null
  __nv_bfloat16* C16;    // [M, ldc16] bf16 or null
  __nv_bfloat16* Ct16;   // transposed [N, ldct] bf16 or null
  const float* bias;     // [N] or null
  const float* row_scale;  // [M] or null: v *= row_scale[row] (applied after relu)
  int64_t ldc, ldc16, ldct;
  float alpha;
  int relu;
  int mode;              // 0 store, 1 C += v (plain RMW), 2 atomicAdd (split-K)
  // optional segment mask: row r is zeroed unless (r % seg_L) < seg_lens[r / seg_L]
  const int* seg_lens;
  int seg_L;
  // optional dropout after relu: injected keep bytes [M, ld_keep] or Philox(seed, stream, r*N+c)
  const uint8_t* keep;
  int64_t ld_keep;
  int philox_dropout;
  float drop_p;
  uint64_t seed, stream;
  int tma_store;         // set by the launcher: fp32 C is written / reduced by TMA from swizzled smem
};

struct EpiStore {
  using Params = EpiStoreParams;
  __device__ __forceinline__ void begin(const Params&, int, int, int, int, int) {}
  __device__ __forceinline__ void end(const Params&, int, int, int, int, int) {}
  // One 32x32 chunk of the accumulator tile: this thread holds row `row` (= row0 + lane), columns
  // col0..col0+31.  The per-row transform runs in registers; the chunk is then staged through
  // shared memory ([32][36] fp32 per warp) so that every global store instruction of the warp
  // covers whole 128-byte lines (4 rows x 128 B for fp32, 8 rows x 64 B for bf16).
  __device__ __forceinline__ void chunk(const Params& p, int row, int col0, float (&v)[32], int M,
                                        int N, float* stage, int lane, int row0) {
    const int ncols = min(32, N - col0);
    if (ncols <= 0 || row0 >= M) return;        // warp-uniform
    transform(p, row, col0, ncols, v, M, N, prefetch_bias(p, col0, N, lane));
    store(p, row, col0, ncols, v, M, N, stage, lane, row0);
  }
  // bias for one 32-column chunk, one value per lane (loaded early, before the accumulator is
  // ready, so that its global-load latency is off the epilogue's critical path)
  __device__ __forceinline__ float prefetch_bias(const Params& p, int col0, int N, int lane) {
    return (p.bias && col0 + lane < N) ? __ldg(p.bias + col0 + lane) : 0.f;
  }
  __device__ __forceinline__ void transform(const Params& p, int row, int col0, int ncols,
                                            float (&v)[32], int M, int N, float bias_lane) {
    // all lanes take part in the shuffles (rows beyond M included)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] * p.alpha + __shfl_sync(0xffffffffu, bias_lane, j);
    if (row < M) {
      float rs = p.row_scale ? p.row_scale[row] : 1.f;
      if (p.seg_lens && (row % p.seg_L) >= p.seg_lens[row / p.seg_L]) rs = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j];
        if (p.relu) x = fmaxf(x, 0.f);
        v[j] = x * rs;
      }
      if (p.keep) {
        const float sc = 1.f / (1.f - p.drop_p);
        const uint8_t* k = p.keep + int64_t(row) * p.ld_keep + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) v[j] = k[j] ? v[j] * sc : 0.f;
      } else if (p.philox_dropout) {
        const float sc = 1.f / (1.f - p.drop_p);
        const uint64_t base = (uint64_t(row) * uint64_t(N) + uint64_t(col0)) >> 2;  // N % 4 == 0
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const uint4 r = Philox::gen(p.seed, p.stream, base + j4);
          v[4 * j4 + 0] = Philox::u01(r.x) >= p.drop_p ? v[4 * j4 + 0] * sc : 0.f;
          v[4 * j4 + 1] = Philox::u01(r.y) >= p.drop_p ? v[4 * j4 + 1] * sc : 0.f;
          v[4 * j4 + 2] = Philox::u01(r.z) >= p.drop_p ? v[4 * j4 + 2] * sc : 0.f;
          v[4 * j4 + 3] = Philox::u01(r.w) >= p.drop_p ? v[4 * j4 + 3] * sc : 0.f;
        }
      }
    }
  }
  __device__ __forceinline__ void store(const Params& p, int row, int col0, int ncols, float (&v)[32],
                                        int M, int N, float* stage, int lane, int row0) {
    if (p.Ct16 && row < M) {
      // for a fixed column the 32 lanes of a warp hold 32 consecutive rows -> coalesced
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) p.Ct16[int64_t(col0 + j) * p.ldct + row] = __float2bfloat16_rn(v[j]);
    }
    if (!p.C && !p.C16) return;
    // stage: thread r writes its row, 16-byte chunk j at position j ^ (r & 7) (conflict-free)
    __syncwarp();
    {
      uint8_t* dst = reinterpret_cast<uint8_t*>(stage) + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(dst + ((j ^ (lane & 7)) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    auto st4 = [&](int r, int c4) -> float4 {
      return *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stage) + r * 128 +
                                              ((c4 ^ (r & 7)) << 4));
    };
    auto st1 = [&](int r, int col) -> float {
      return stage[r * 32 + ((((col >> 2) ^ (r & 7)) << 2) | (col & 3))];
    };
    const int rows_here = min(32, M - row0);
    if (p.C) {
      float* base = p.C + int64_t(row0) * p.ldc + col0;
      const bool vec = (ncols == 32) && ((p.ldc & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
      if (p.mode == 2) {
        // split-K reduction: one row per instruction, 32 consecutive addresses per warp
        for (int r = 0; r < rows_here; ++r)
          if (lane < ncols) atomicAdd(base + int64_t(r) * p.ldc + lane, st1(r, lane));
      } else if (vec) {
        const int rsub = lane >> 3, c4 = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + rsub;
          if (r < rows_here) {
            float4 o = st4(r, c4);
            float4* d = reinterpret_cast<float4*>(base + int64_t(r) * p.ldc) + c4;
            if (p.mode == 1) {
              const float4 old = *d;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *d = o;
          }
        }
      } else {
        for (int r = 0; r < rows_here; ++r)
          if (lane < ncols) {
            float* d = base + int64_t(r) * p.ldc + lane;
            const float x = st1(r, lane);
            *d = (p.mode == 1) ? *d + x : x;
          }
      }
    }
    if (p.C16) {
      __nv_bfloat16* base = p.C16 + int64_t(row0) * p.ldc16 + col0;
      const bool vec = (ncols == 32) && ((p.ldc16 & 7) == 0) &&
                       ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
      if (vec) {
        const int rsub = lane >> 2, c8 = lane & 3;   // 8 rows x (4 lanes x 8 bf16) per instruction
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + rsub;
          if (r < rows_here) {
            const float4 a = st4(r, 2 * c8);
            const float4 b = st4(r, 2 * c8 + 1);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&h0);
            o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2);
            o.w = *reinterpret_cast<uint32_t*>(&h3);
            *(reinterpret_cast<uint4*>(base + int64_t(r) * p.ldc16) + c8) = o;
          }
        }
      } else {
        for (int r = 0; r < rows_here; ++r)
          if (lane < ncols) base[int64_t(r) * p.ldc16 + lane] = __float2bfloat16_rn(st1(r, lane));
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// main-loop roles, shared by every kernel built on this pipeline (gemm_tc_kernel and the fused
// logit + sampling kernel of logit_sample.cuh).  Tiles are dealt round-robin, m fastest:
//   t = blockIdx.x + i * gridDim.x ;  m_blk = t % num_m ;  n_blk = (t / num_m) % num_n ;  ks = ...
// ------------------------------------------------------------------------------------------
// one elected thread: TMA loads of the A / B k-blocks into the smem ring
template <class Cfg, int BN, int AMAJ, int BMAJ>
__device__ __forceinline__ void gemm_producer_role(const CUtensorMap* tmA, const CUtensorMap* tmB,
                                                   uint8_t* sA, uint8_t* sB, uint64_t* full_bar,
                                                   uint64_t* empty_bar, int num_m, int num_n,
                                                   int num_tiles, int nkb, int kb_per_split) {
  constexpr int STAGES = Cfg::STAGES;
  int st = 0;
  uint32_t ph = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int m_blk = t % num_m;
    const int rest = t / num_m;
    const int n_blk = rest % num_n;
    const int ks = rest / num_n;
    const int kb0 = ks * kb_per_split;
    const int kb1 = min(nkb, kb0 + kb_per_split);
    const int m0 = m_blk * GEMM_BM, n0 = n_blk * BN;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty_bar[st], ph ^ 1);
      mbar_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
      uint8_t* a_dst = sA + st * Cfg::A_BYTES;
      uint8_t* b_dst = sB + st * Cfg::B_BYTES;
      const int k0 = kb * Cfg::BK;
      if constexpr (AMAJ == 0) {
        tma_load_2d(a_dst, tmA, &full_bar[st], k0, m0);
      } else {
#pragma unroll
        for (int j = 0; j < GEMM_BM / Cfg::MN_ATOM; ++j)
          tma_load_2d(a_dst + j * (Cfg::BK * 128), tmA, &full_bar[st], m0 + j * Cfg::MN_ATOM, k0);
      }
      if constexpr (BMAJ == 0) {
        tma_load_2d(b_dst, tmB, &full_bar[st], k0, n0);
      } else {
#pragma unroll
        for (int j = 0; j < BN / Cfg::MN_ATOM; ++j)
          tma_load_2d(b_dst + j * (Cfg::BK * 128), tmB, &full_bar[st], n0 + j * Cfg::MN_ATOM, k0);
      }
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  }
}

// one elected thread: tcgen05.mma over the ring into the double-buffered TMEM accumulator
template <int KIND, class Cfg, int BN, int AMAJ, int BMAJ>
__device__ __forceinline__ void gemm_mma_role(uint8_t* sA, uint8_t* sB, uint64_t* full_bar,
                                              uint64_t* empty_bar, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t tmem_base, int num_m,
                                              int num_n, int num_tiles, int nkb, int kb_per_split) {
  constexpr int STAGES = Cfg::STAGES;
  constexpr uint32_t idesc = make_idesc<KIND>(GEMM_BM, BN, AMAJ, BMAJ);
  int st = 0;
  uint32_t ph = 0;
  int as = 0;
  uint32_t aph = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int ks = (t / num_m) / num_n;
    const int kb0 = ks * kb_per_split;
    const int kb1 = min(nkb, kb0 + kb_per_split);
    if (kb0 >= kb1) continue;
    mbar_wait(&tempty_bar[as], aph ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + as * BN;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full_bar[st], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA + st * Cfg::A_BYTES);
      const uint32_t b_addr = smem_u32(sB + st * Cfg::B_BYTES);
#pragma unroll
      for (int k = 0; k < Cfg::BK / Cfg::UK; ++k) {
        const uint64_t da = (AMAJ == 0)
                                ? make_smem_desc(a_addr + k * 32, 16, 1024)
                                : make_smem_desc(a_addr + k * Cfg::UK * 128, Cfg::BK * 128, 1024);
        const uint64_t db = (BMAJ == 0)
                                ? make_smem_desc(b_addr + k * 32, 16, 1024)
                                : make_smem_desc(b_addr + k * Cfg::UK * 128, Cfg::BK * 128, 1024);
        umma_ss<KIND>(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[st]);
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
    umma_commit(&tfull_bar[as]);
    if (++as == 2) { as = 0; aph ^= 1; }
  }
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
// LEAN = true: the epilogue is only alpha * acc + bias -> fp32 C through TMA store / reduce-add, as a
// rolled loop over the 32-column chunks.  The generic epilogue (relu, dropout, masks, bf16 /
// transposed outputs, three store modes) unrolls to ~200 KB of SASS per instantiation; the ~100
// per-step GEMMs of the decode loop run ONE tile per CTA, so every instruction they execute is an
// instruction-cache miss (ncu: stall_no_instruction 8.0 per issued instruction on the generic
// kernel).  The lean variant keeps their code to a few KB.
template <int KIND, int BN, int AMAJ, int BMAJ, class Epi, bool LEAN = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)   // 10 warps are allocated as 12: 168 regs/thread
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int split_k,
               typename Epi::Params ep) {
  using Cfg = GemmCfg<KIND, BN>;
  constexpr int STAGES = Cfg::STAGES;
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
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES + GEMM_SMEM_EXTRA;   // 1024-aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int nkb = (K + Cfg::BK - 1) / Cfg::BK;
  const int kb_per_split = (nkb + split_k - 1) / split_k;
  const int num_tiles = num_m * num_n * split_k;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ep.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // setup above overlapped the previous kernel's tail; from here on global memory is touched
  pdl_wait();

  if (warp == 0) {
    if (lane == 0)
      gemm_producer_role<Cfg, BN, AMAJ, BMAJ>(&tmA, &tmB, sA, sB, full_bar, empty_bar, num_m, num_n,
                                              num_tiles, nkb, kb_per_split);
  } else if (warp == 1) {
    if (lane == 0)
      gemm_mma_role<KIND, Cfg, BN, AMAJ, BMAJ>(sA, sB, full_bar, empty_bar, tfull_bar, tempty_bar,
                                               tmem_base, num_m, num_n, num_tiles, nkb, kb_per_split);
  } else {
    // ===================== epilogue (warps 2..9: two warps per TMEM lane quarter) =====================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the 32-column chunks this warp handles
    Epi epi;
    int as = 0;
    uint32_t aph = 0;
    int ebuf = 0;   // staging buffer of the next TMA store; alternates across chunks AND tiles
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % num_m;
      const int rest = t / num_m;
      const int n_blk = rest % num_n;
      const int ks = rest / num_n;
      const int kb0 = ks * kb_per_split;
      const int kb1 = min(nkb, kb0 + kb_per_split);
      if (kb0 >= kb1) continue;
      const int row = m_blk * GEMM_BM + q * 32 + lane;
      const int n0 = n_blk * BN;
      if constexpr (LEAN) {
        constexpr int NCL = BN / 32 / 2;
        const int row0 = m_blk * GEMM_BM + q * 32;
        const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
        uint8_t* wstage = epi_stage + (warp - 2) * Cfg::EPI_BUFS * 4096;
        float bias_lane = (ep.bias && n0 + half * 32 + lane < N) ? __ldg(ep.bias + n0 + half * 32 + lane) : 0.f;
        mbar_wait(&tfull_bar[as], aph);
        tc_fence_after();
#pragma unroll 1
        for (int i = 0; i < NCL; ++i) {
          const int col0 = n0 + (half + 2 * i) * 32;
          float v[32];
          tmem_ld32(t_addr + (half + 2 * i) * 32, v);
          // next chunk's bias while the accumulator columns are in flight
          const int coln = col0 + 64 + lane;
          const float bias_next = (ep.bias && i + 1 < NCL && coln < N) ? __ldg(ep.bias + coln) : 0.f;
          tmem_ld_wait();
          if (i + 1 == NCL) {
            tc_fence_before();
            mbar_arrive(&tempty_bar[as]);
          }
          if (col0 < N && row0 < M) {            // warp-uniform
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] * ep.alpha + __shfl_sync(0xffffffffu, bias_lane, j);
            uint8_t* buf = wstage + (Cfg::EPI_BUFS == 2 ? ebuf * 4096 : 0);
            ebuf ^= 1;
            if (lane == 0) bulk_wait_read<Cfg::EPI_BUFS - 1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (ep.mode == 0) tma_store_2d(&tmC, buf, col0, row0);
              else tma_reduce_add_2d(&tmC, buf, col0, row0);
              bulk_commit();
            }
          }
          bias_lane = bias_next;
        }
        if (++as == 2) { as = 0; aph ^= 1; }
        continue;
      }
      float bias_pf[BN / 64];
#pragma unroll
      for (int i = 0; i < BN / 64; ++i) bias_pf[i] = epi.prefetch_bias(ep, n0 + (half + 2 * i) * 32, N, lane);
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      epi.begin(ep, row, n0, M, N, ks);
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
      uint8_t* wstage = epi_stage + (warp - 2) * Cfg::EPI_BUFS * 4096;
      const int row0 = m_blk * GEMM_BM + q * 32;
      constexpr int NC = BN / 32 / 2;    // chunks per warp: c = half, half + 2, ...
      float v[2][32];
      tmem_ld32(t_addr + half * 32, v[0]);
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        tmem_ld_wait();
        if (i + 1 < NC) tmem_ld32(t_addr + (half + 2 * (i + 1)) * 32, v[(i + 1) & 1]);
        else {
          // this thread has read all of its accumulator columns: hand the TMEM buffer back to the
          // MMA warp now, the stores of this chunk overlap the next tile's MMAs
          tc_fence_before();
          mbar_arrive(&tempty_bar[as]);
        }
        const int col0 = n0 + (half + 2 * i) * 32;
        if (ep.tma_store) {
          if (col0 < N && row0 < M) {            // warp-uniform
            epi.transform(ep, row, col0, min(32, N - col0), v[i & 1], M, N, bias_pf[i]);
            uint8_t* buf = wstage + (Cfg::EPI_BUFS == 2 ? ebuf * 4096 : 0);
            ebuf ^= 1;
            // the bulk store that last read this buffer must have drained
            if (lane == 0) bulk_wait_read<Cfg::EPI_BUFS - 1>();
            __syncwarp();
            // row `lane` -> 128 B, 16-byte chunk j stored at (j ^ (lane & 7)): SWIZZLE_128B layout
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[i & 1][4 * j], v[i & 1][4 * j + 1], v[i & 1][4 * j + 2],
                              v[i & 1][4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (ep.mode == 0) tma_store_2d(&tmC, buf, col0, row0);
              else tma_reduce_add_2d(&tmC, buf, col0, row0);
              bulk_commit();
            }
          }
        } else {
          if (col0 < N && row0 < M) {
            epi.transform(ep, row, col0, min(32, N - col0), v[i & 1], M, N, bias_pf[i]);
            epi.store(ep, row, col0, min(32, N - col0), v[i & 1], M, N,
                      reinterpret_cast<float*>(wstage), lane, row0);
          }
        }
      }
      epi.end(ep, row, n0, M, N, n_blk);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }
  if (warp >= 2 && lane == 0 && ep.tma_store) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// Encode a 2D row-major [rows, cols] matrix (leading dimension ld elements) as a TMA tensor map
// with a [box_rows, box_cols] box and 128B swizzle. elem_bytes 2 -> bf16, 4 -> fp32.
int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, int64_t rows, int64_t cols,
                   int64_t ld, int box_rows, int box_cols);

template <int KIND, int BN, int AMAJ, int BMAJ, class Epi>
int launch_gemm_tc(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                   int split_k, const typename Epi::Params& ep_in, cudaStream_t stream,
                   int max_ctas = 0) {
  typename Epi::Params ep = ep_in;
  using Cfg = GemmCfg<KIND, BN>;
  CC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  CC_REQUIRE(split_k >= 1, "gemm: split_k must be >= 1");
  CUtensorMap tmA, tmB;
  int rc;
  if (AMAJ == 0)
    rc = encode_tmap_2d(&tmA, A, Cfg::EB, M, K, lda, GEMM_BM, Cfg::BK);
  else
    rc = encode_tmap_2d(&tmA, A, Cfg::EB, K, M, lda, Cfg::BK, Cfg::MN_ATOM);
  if (rc) return rc;
  if (BMAJ == 0)
    rc = encode_tmap_2d(&tmB, B, Cfg::EB, N, K, ldb, BN, Cfg::BK);
  else
    rc = encode_tmap_2d(&tmB, B, Cfg::EB, K, N, ldb, Cfg::BK, Cfg::MN_ATOM);
  if (rc) return rc;

  // fp32-only output with a TMA-compatible pitch: the epilogue stores / reduces through TMA
  CUtensorMap tmC = tmA;
  ep.tma_store = 0;
  if (ep.C && !ep.C16 && !ep.Ct16 && (ep.ldc % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(ep.C) & 15) == 0) {
    rc = encode_tmap_2d(&tmC, ep.C, 4, M, N, ep.ldc, 32, 32);
    if (rc) return rc;
    ep.tma_store = 1;
  }
  // plain "alpha * acc + bias -> fp32 C by TMA" problems take the small-code kernel
  const bool lean = ep.tma_store && !ep.relu && !ep.row_scale && !ep.seg_lens && !ep.keep &&
                    !ep.philox_dropout;
  auto kern = lean ? gemm_tc_kernel<KIND, BN, AMAJ, BMAJ, Epi, true>
                   : gemm_tc_kernel<KIND, BN, AMAJ, BMAJ, Epi, false>;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES))) return rc;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM, num_n = (N + BN - 1) / BN;
  const int nkb = (K + Cfg::BK - 1) / Cfg::BK;
  if (split_k > nkb) split_k = nkb;
  const int tiles = num_m * num_n * split_k;
  int grid = num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (tiles < grid) grid = tiles;
  CC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, tmC,
                           M, N, K, split_k, ep));
  CC_LAUNCH_CHECK_K(PROF_GEMM, stream, 2.0 * double(M) * double(N) * double(K),
                    double(Cfg::EB) * (double(M) * K + double(N) * K));
  return CC_OK;
}

// Runtime-dispatched GEMM with the generic store epilogue (defined in gemm_api.cu).
int gemm_run(int kind, int a_major, int b_major, const void* A, int64_t lda, const void* B,
             int64_t ldb, int M, int N, int K, int split_k, int tile_n, const EpiStoreParams& ep,
             cudaStream_t s);

}  // namespace coopcap

struct coopcap_gemm_args;
namespace coopcap {
// Full argument-struct entry (tcgen05 or the SIMT cross-check backend), defined in gemm_api.cu.
int gemm_store(const coopcap_gemm_args* a, cudaStream_t s);
}  // namespace coopcap
