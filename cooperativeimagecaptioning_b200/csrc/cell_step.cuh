// Recurrent-cell steps as ONE kernel: the step's tcgen05 GEMM with the cell's pointwise update as
// its epilogue.
//
//   GRU  (VSEFCModel.py:108-112, torch.nn.GRU gate order r, z, n):
//        gh = h_{t-1} . W_hh^T + b_hh ;  r, z, n, h_t  -- replaces gemm + gru_fwd_kernel
//   LSTM (AttModel.py:515-531, the att2in2 maxout cell):
//        u = att_res . W_a2c^T + b_a2c ; c_t, h_t, dropout(h_t) -- replaces gemm + lstm_fwd_kernel
//
// A cell needs, for one hidden unit j, NG columns of the GEMM that lie M (or R) apart in the weight
// matrix (gates r|z|n of unit j; the two maxout halves of unit j).  The producer therefore builds
// the B tile out of NG boxes of UNITS weight rows each -- rows [g * gate_stride + j0, +UNITS) --
// stacked in shared memory, which for the K-major 128B-swizzled layout is byte-identical to one
// NG*UNITS-row box: no permuted copy of the weights exists.  The accumulator tile is then
// [128 rows] x [gate 0: UNITS | gate 1: UNITS | ...] in TMEM.
//
// Epilogue: tcgen05.ld hands a thread one ROW of the tile, but the cell's other operands (gi / the
// saved gate pre-activations, c_{t-1}, h_{t-1}) and its outputs are row-major in global memory, so a
// thread-per-row access pattern would touch 32 lines per instruction.  Each warp passes its
// [32 rows] x [16 units] x NG sub-block through a swizzled shared-memory stage and does the
// pointwise work in the transposed assignment (8 rows x 4 lanes, 4 units per lane): every global
// instruction of the warp covers whole 64-byte row segments.
//
// The arithmetic (operation order, expf / tanhf, Philox counters of the dropout mask) is the one of
// gru_fwd_kernel / lstm_fwd_kernel: the fused step is bit-identical to the two-kernel path
// (tests/test_gpu_cell_fuse.py).
//
// MEASURED AND NOT THE DEFAULT (COOPCAP_CELL_FUSE=1 opts in).  Same-box A/B on the 1024-row Gumbel
// joint step (gpurun_out/s6_bench*.json, r2): 260 instead of 293 launches per step, but 5.603 ms
// against 5.583 ms for the two-kernel path.  The cell update is ~150 instructions per hidden unit
// (IEEE divide, expf, tanhf) and moves 300 KB per 128-row tile; as an epilogue it runs on 8 warps of
// 128 SMs after the main loop has finished, while the stand-alone pointwise kernel spreads the same
// work over 64 warps on each of 148 SMs and, launched with PDL, hides its launch latency behind the
// GEMM's tail.  The GEMM class grows by 6.7 us per fused launch, the pointwise classes shrink by as
// much.  Requesting the cell's operands before the accumulator wait (below) recovered 0.05 ms of an
// initial 0.07 ms loss.
#pragma once
#include "gemm.cuh"
#include "speaker_kernels.cuh"

namespace coopcap {

constexpr int CELL_EPI_WARPS = 8;
constexpr int CELL_THREADS = (2 + CELL_EPI_WARPS) * 32;
constexpr int CELL_SUB = 16;              // units per staged sub-block

template <int NG_, int UNITS_>
struct CellCfg {                           // bf16 operands, both K-major
  static constexpr int NG = NG_, UNITS = UNITS_;
  static constexpr int BN = NG * UNITS;
  static constexpr int EB = 2, BK = 64, UK = 16, MN_ATOM = 64;
  static constexpr int A_BYTES = GEMM_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int WARP_STAGE_BYTES = NG * 32 * CELL_SUB * 4;
  static constexpr int EPI_STAGE_BYTES = CELL_EPI_WARPS * WARP_STAGE_BYTES;
  static constexpr int STAGES_RAW =
      (GEMM_SMEM_TOTAL - GEMM_SMEM_EXTRA - 1024 - EPI_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_SMEM_EXTRA + EPI_STAGE_BYTES + 1024;
  static_assert(UNITS % (2 * CELL_SUB) == 0, "each epilogue warp half takes whole 16-unit sub-blocks");
  static_assert(BN % 16 == 0 && BN <= 256 && 2 * BN <= 512, "tile width");
  static_assert(STAGES >= 3, "pipeline too shallow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};
using GruCfg = CellCfg<3, 64>;             // 128 x (3 x 64): 16 column tiles for M = 1024
using LstmCfg = CellCfg<2, 32>;            // 128 x (2 x 32): 16 column tiles for R = 512

struct GruStepParams {
  const float* b_hh;        // [3M]
  const float* gi;          // [B, 3M] input pre-activations of this step (bias b_ih included)
  const float* h_prev;      // [B, M]
  const int* len;           // [B]
  float* gates;             // [B, 4M]  r | z | n | gh_n (what gru_bwd_kernel reads)
  float* h_next;            // [B, M]
  __nv_bfloat16* h_next16;  // [B, M]   next step's GEMM operand
  int t;
};

struct LstmStepParams {
  const float* b_a2c;       // [2R]
  const float* s;           // [B, lds] gate pre-activations i | f | o | maxout 1 | maxout 2 (| att_h)
  int64_t lds;
  const float* c_prev;      // [B, R]
  float* u;                 // [B, 2R]  a2c(att_res) + bias, kept for backward
  float* c_next;            // [B, R]
  __nv_bfloat16* h_dst;     // h_t into the next step's [x | h] operand
  int64_t ld_h;
  __nv_bfloat16* out16;     // [B, R]   dropout(h_t): the logit GEMM's operand
  const uint8_t* keep;      // injected keep mask [B, R] or null (Philox)
  uint64_t seed, stream;
  float drop_p;
};

// one elected thread: A as in gemm_producer_role, B as NG stacked boxes of UNITS weight rows
template <class Cfg>
__device__ __forceinline__ void cell_producer_role(const CUtensorMap* tmA, const CUtensorMap* tmB,
                                                   uint8_t* sA, uint8_t* sB, uint64_t* full_bar,
                                                   uint64_t* empty_bar, int num_m, int num_tiles,
                                                   int nkb, int gate_stride) {
  constexpr int STAGES = Cfg::STAGES;
  int st = 0;
  uint32_t ph = 0;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int m0 = (t % num_m) * GEMM_BM;
    const int j0 = (t / num_m) * Cfg::UNITS;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&empty_bar[st], ph ^ 1);
      mbar_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
      const int k0 = kb * Cfg::BK;
      tma_load_2d(sA + st * Cfg::A_BYTES, tmA, &full_bar[st], k0, m0);
#pragma unroll
      for (int g = 0; g < Cfg::NG; ++g)
        tma_load_2d(sB + st * Cfg::B_BYTES + g * (Cfg::UNITS * 128), tmB, &full_bar[st], k0,
                    g * gate_stride + j0);
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  }
}

// staging layout of one gate's [32 rows][16 units] fp32 sub-block: row r is 64 bytes, its 16-byte
// piece j sits at piece j ^ ((r >> 1) & 3) -- conflict-free for the row-per-lane writes and for the
// 8-rows-by-4-lanes reads
__device__ __forceinline__ uint32_t cell_stage_off(int r, int piece) {
  return uint32_t(r) * 64u + (uint32_t(piece ^ ((r >> 1) & 3)) << 4);
}

struct GruCell {
  using Cfg = GruCfg;
  using Params = GruStepParams;
  static constexpr int PF = 4;             // work items (8 rows x 16 units each) whose operands are in flight
  struct Ops { float4 ir, iz, in, hp; int active; };
  struct Bias { float4 r, z, n; };
  static __device__ __forceinline__ Bias bias(const Params& p, int j, int Mh) {
    Bias b;
    b.r = __ldg(reinterpret_cast<const float4*>(p.b_hh + j));
    b.z = __ldg(reinterpret_cast<const float4*>(p.b_hh + Mh + j));
    b.n = __ldg(reinterpret_cast<const float4*>(p.b_hh + 2 * Mh + j));
    return b;
  }
  // everything the update of row b, units j..j+3 reads besides the accumulators
  static __device__ __forceinline__ Ops load(const Params& p, int b, int j, int Mh) {
    Ops o;
    const float* gib = p.gi + int64_t(b) * 3 * Mh;
    o.ir = *reinterpret_cast<const float4*>(gib + j);
    o.iz = *reinterpret_cast<const float4*>(gib + Mh + j);
    o.in = *reinterpret_cast<const float4*>(gib + 2 * Mh + j);
    o.hp = *reinterpret_cast<const float4*>(p.h_prev + int64_t(b) * Mh + j);
    o.active = p.t < p.len[b];
    return o;
  }
  // g[k] = gate k's four accumulators
  static __device__ __forceinline__ void finish(const Params& p, int b, int j, int Mh,
                                                const float4 (&g)[3], const Ops& op, const Bias& bb) {
    const float a_ir[4] = {op.ir.x, op.ir.y, op.ir.z, op.ir.w}, a_iz[4] = {op.iz.x, op.iz.y, op.iz.z, op.iz.w};
    const float a_in[4] = {op.in.x, op.in.y, op.in.z, op.in.w};
    const float a_hr[4] = {g[0].x + bb.r.x, g[0].y + bb.r.y, g[0].z + bb.r.z, g[0].w + bb.r.w};
    const float a_hz[4] = {g[1].x + bb.z.x, g[1].y + bb.z.y, g[1].z + bb.z.z, g[1].w + bb.z.w};
    const float a_hn[4] = {g[2].x + bb.n.x, g[2].y + bb.n.y, g[2].z + bb.n.z, g[2].w + bb.n.w};
    const float a_hp[4] = {op.hp.x, op.hp.y, op.hp.z, op.hp.w};
    const bool active = op.active != 0;
    float r[4], z[4], n[4], h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r[q] = 1.f / (1.f + expf(-(a_ir[q] + a_hr[q])));
      z[q] = 1.f / (1.f + expf(-(a_iz[q] + a_hz[q])));
      n[q] = tanhf(a_in[q] + r[q] * a_hn[q]);
      h[q] = active ? (1.f - z[q]) * n[q] + z[q] * a_hp[q] : a_hp[q];
    }
    float* go = p.gates + int64_t(b) * 4 * Mh;
    *reinterpret_cast<float4*>(go + j) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(go + Mh + j) = make_float4(z[0], z[1], z[2], z[3]);
    *reinterpret_cast<float4*>(go + 2 * Mh + j) = make_float4(n[0], n[1], n[2], n[3]);
    *reinterpret_cast<float4*>(go + 3 * Mh + j) = make_float4(a_hn[0], a_hn[1], a_hn[2], a_hn[3]);
    *reinterpret_cast<float4*>(p.h_next + int64_t(b) * Mh + j) = make_float4(h[0], h[1], h[2], h[3]);
    __nv_bfloat162 a = __floats2bfloat162_rn(h[0], h[1]), c = __floats2bfloat162_rn(h[2], h[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&c);
    *reinterpret_cast<uint2*>(p.h_next16 + int64_t(b) * Mh + j) = o;
  }
};

struct LstmCell {
  using Cfg = LstmCfg;
  using Params = LstmStepParams;
  static constexpr int PF = 2;
  struct Ops { float4 si, sf, so, s1, s2, cp; };
  struct Bias { float4 b1, b2; };
  static __device__ __forceinline__ Bias bias(const Params& p, int j, int R) {
    Bias b;
    b.b1 = __ldg(reinterpret_cast<const float4*>(p.b_a2c + j));
    b.b2 = __ldg(reinterpret_cast<const float4*>(p.b_a2c + R + j));
    return b;
  }
  static __device__ __forceinline__ Ops load(const Params& p, int b, int j, int R) {
    Ops o;
    const float* sr = p.s + int64_t(b) * p.lds;
    o.si = *reinterpret_cast<const float4*>(sr + j);
    o.sf = *reinterpret_cast<const float4*>(sr + R + j);
    o.so = *reinterpret_cast<const float4*>(sr + 2 * R + j);
    o.s1 = *reinterpret_cast<const float4*>(sr + 3 * R + j);
    o.s2 = *reinterpret_cast<const float4*>(sr + 4 * R + j);
    o.cp = *reinterpret_cast<const float4*>(p.c_prev + int64_t(b) * R + j);
    return o;
  }
  static __device__ __forceinline__ void finish(const Params& p, int b, int j, int R,
                                                const float4 (&g)[2], const Ops& op, const Bias& bb) {
    const float4 u1 = make_float4(g[0].x + bb.b1.x, g[0].y + bb.b1.y, g[0].z + bb.b1.z, g[0].w + bb.b1.w);
    const float4 u2 = make_float4(g[1].x + bb.b2.x, g[1].y + bb.b2.y, g[1].z + bb.b2.z, g[1].w + bb.b2.w);
    *reinterpret_cast<float4*>(p.u + int64_t(b) * 2 * R + j) = u1;
    *reinterpret_cast<float4*>(p.u + int64_t(b) * 2 * R + R + j) = u2;
    const float ai[4] = {op.si.x, op.si.y, op.si.z, op.si.w}, af[4] = {op.sf.x, op.sf.y, op.sf.z, op.sf.w};
    const float ao[4] = {op.so.x, op.so.y, op.so.z, op.so.w};
    const float a1[4] = {op.s1.x + u1.x, op.s1.y + u1.y, op.s1.z + u1.z, op.s1.w + u1.w};
    const float a2[4] = {op.s2.x + u2.x, op.s2.y + u2.y, op.s2.z + u2.z, op.s2.w + u2.w};
    const float ac[4] = {op.cp.x, op.cp.y, op.cp.z, op.cp.w};
    float cn[4], h[4], o[4];
    bool k[4] = {true, true, true, true};
    const float sc = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    if (p.drop_p > 0.f) keep4(p.keep, (int64_t(b) * R + j) >> 2, p.seed, p.stream, p.drop_p, k);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float ig = 1.f / (1.f + expf(-ai[q]));
      const float fg = 1.f / (1.f + expf(-af[q]));
      const float og = 1.f / (1.f + expf(-ao[q]));
      const float gq = fmaxf(a1[q], a2[q]);
      cn[q] = fg * ac[q] + ig * gq;
      h[q] = og * tanhf(cn[q]);
      o[q] = k[q] ? h[q] * sc : 0.f;
    }
    *reinterpret_cast<float4*>(p.c_next + int64_t(b) * R + j) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    {
      __nv_bfloat162 a = __floats2bfloat162_rn(h[0], h[1]), b2 = __floats2bfloat162_rn(h[2], h[3]);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&a);
      w.y = *reinterpret_cast<uint32_t*>(&b2);
      *reinterpret_cast<uint2*>(p.h_dst + int64_t(b) * p.ld_h + j) = w;
    }
    {
      __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b2 = __floats2bfloat162_rn(o[2], o[3]);
      uint2 w;
      w.x = *reinterpret_cast<uint32_t*>(&a);
      w.y = *reinterpret_cast<uint32_t*>(&b2);
      *reinterpret_cast<uint2*>(p.out16 + int64_t(b) * R + j) = w;
    }
  }
};

// C[M, NG * H] = A[M, K] . W[NG * H, K]^T with the cell update as epilogue; H = hidden size
template <class Cell>
__global__ void __launch_bounds__(CELL_THREADS, 1)
cell_step_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 int M, int H, int K, typename Cell::Params p) {
  using Cfg = typename Cell::Cfg;
  constexpr int BN = Cfg::BN, NG = Cfg::NG, UNITS = Cfg::UNITS;
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
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES + GEMM_SMEM_EXTRA;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = H / UNITS;
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
      mbar_init(&tempty_bar[s], CELL_EPI_WARPS * 32);
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
      cell_producer_role<Cfg>(&tmA, &tmB, sA, sB, full_bar, empty_bar, num_m, num_tiles, nkb, H);
  } else if (warp == 1) {
    if (lane == 0)
      gemm_mma_role<0, Cfg, BN, 0, 0>(sA, sB, full_bar, empty_bar, tfull_bar, tempty_bar, tmem_base,
                                      num_m, num_n, num_tiles, nkb, nkb);
  } else {
    const int q = warp & 3;              // TMEM lane quarter: rows q*32 .. q*32+31 of the tile
    const int half = (warp - 2) >> 2;    // which half of the tile's units
    constexpr int NSUB = UNITS / 2 / CELL_SUB;
    uint8_t* wstage = epi_stage + (warp - 2) * Cfg::WARP_STAGE_BYTES;
    const int rsub = lane >> 2, c4 = lane & 3;
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % num_m, n_blk = t / num_m;
      const int row0 = m_blk * GEMM_BM + q * 32;
      // Work items of this lane: k = c * 4 + i -> sub-block c, rows i*8 + rsub, 4 units.  The operands
      // of the first PF items are requested BEFORE the accumulator is waited for (they do not depend
      // on it: their L2 / DRAM latency hides behind the main loop); each finished item requests
      // item k + PF.
      constexpr int ITEMS = 4 * NSUB, PF = Cell::PF;
      auto item_row = [&](int k) { return row0 + (k & 3) * 8 + rsub; };
      auto item_unit = [&](int k) { return n_blk * UNITS + half * (UNITS / 2) + (k >> 2) * CELL_SUB + c4 * 4; };
      typename Cell::Ops ops[PF];
      typename Cell::Bias bias = Cell::bias(p, item_unit(0), H);
#pragma unroll
      for (int k = 0; k < PF; ++k)
        if (k < ITEMS && item_row(k) < M) ops[k] = Cell::load(p, item_row(k), item_unit(k), H);
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;
#pragma unroll
      for (int c = 0; c < NSUB; ++c) {
        const int uo = half * (UNITS / 2) + c * CELL_SUB;      // unit offset within the tile
        {
          // gate by gate, the next gate's columns in flight while this one is staged (two batches of
          // accumulators live instead of NG)
          float x[2][CELL_SUB];
          tmem_ld16(t_addr + uo, x[0]);
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            tmem_ld_wait();
            if (g + 1 < NG) tmem_ld16(t_addr + (g + 1) * UNITS + uo, x[(g + 1) & 1]);
            else if (c == NSUB - 1) {      // all of this thread's columns are out of TMEM
              tc_fence_before();
              mbar_arrive(&tempty_bar[as]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(wstage + g * (32 * CELL_SUB * 4) + cell_stage_off(lane, j)) =
                  make_float4(x[g & 1][4 * j], x[g & 1][4 * j + 1], x[g & 1][4 * j + 2], x[g & 1][4 * j + 3]);
          }
        }
        __syncwarp();
        // the next sub-block's bias slice is requested only now: the accumulator batch above is dead,
        // so the registers exist (all NSUB slices up front spilled)
        typename Cell::Bias bias_next = bias;
        if (c + 1 < NSUB) bias_next = Cell::bias(p, item_unit(4 * (c + 1)), H);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = c * 4 + i;
          const int r = i * 8 + rsub;
          const int b = row0 + r;
          if (b < M) {
            float4 gv[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g)
              gv[g] = *reinterpret_cast<const float4*>(wstage + g * (32 * CELL_SUB * 4) +
                                                       cell_stage_off(r, c4));
            Cell::finish(p, b, item_unit(k), H, gv, ops[k % PF], bias);
          }
          if (k + PF < ITEMS && item_row(k + PF) < M)
            ops[k % PF] = Cell::load(p, item_row(k + PF), item_unit(k + PF), H);
        }
        bias = bias_next;
        __syncwarp();
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

inline bool cell_fuse_enabled() {
  static const bool on = [] { const char* e = getenv("COOPCAP_CELL_FUSE"); return e && e[0] == '1'; }();
  return on;
}

// A: [M, K] bf16 (leading dimension lda), W: [NG * H, K] bf16 row-major.  Returns CC_OK after the
// launch; the caller checks `cell_step_ok` first.
template <class Cell>
inline bool cell_step_ok(int H, int K) {
  return cell_fuse_enabled() && H % Cell::Cfg::UNITS == 0 && K % 8 == 0 && H % 4 == 0;
}

template <class Cell>
int launch_cell_step(const void* A, int64_t lda, const void* W, int M, int H, int K,
                     const typename Cell::Params& p, cudaStream_t s) {
  using Cfg = typename Cell::Cfg;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, A, 2, M, K, lda, GEMM_BM, Cfg::BK))) return rc;
  if ((rc = encode_tmap_2d(&tmB, W, 2, int64_t(Cfg::NG) * H, K, K, Cfg::UNITS, Cfg::BK))) return rc;
  auto kern = cell_step_kernel<Cell>;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES))) return rc;
  const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * (H / Cfg::UNITS);
  const int grid = std::min(num_sms(), tiles);
  CC_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(CELL_THREADS), size_t(Cfg::SMEM_BYTES), s, tmA, tmB, M,
                           H, K, p));
  prof_mark(PROF_GEMM, s, 2.0 * double(M) * double(Cfg::NG) * H * K,
            2.0 * (double(M) * K + double(Cfg::NG) * H * K));
  return CC_OK;
}

}  // namespace coopcap
