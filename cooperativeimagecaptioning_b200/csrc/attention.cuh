// Additive attention over packed region features (AttModel.py:465-489), forward / per-step
// backward / deferred region-gradient accumulation.
//
// All three kernels are streams over the packed region tensors:
//   forward : one pass over (p_att, att_e) with an online softmax (running max / sum / weighted sum)
//   backward: pass 1 over att_e (dw_l = <d_att_res, att_e_l>), softmax backward, pass 2 over p_att
//             (d_att_h = alpha * sum_l de_l (1 - tanh^2))
//   deferred: one pass over p_att for all steps at once (d_p_att, d_att_e, alpha / bias partials),
//             one CTA per row, 8-region chunks through a TMA bulk-copy ring
// forward / backward run once per decode step: persistent CTAs, two warps per row, per-warp
// cp.async rings (see the v4 note below).
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
// v4: one persistent CTA per SM, 16 warps, TWO WARPS PER ROW (each takes half of the row's
// regions).  Every warp owns a private 3-stage cp.async ring (16-byte pieces per lane, 12 KB) and
// refills a stage right after consuming it, so ~190 KB of region data stay in flight per SM with
// no block-wide barrier anywhere; the two half-row states meet once through a 64-thread named
// barrier.  (Measured on B200: per-warp cp.async.bulk copies of 2 KB are issue-rate bound at
// 2.2 TB/s, 4 KB at 3.6 TB/s; plain 16-byte streams reach 4.4-4.6 TB/s on this 117 MB read.)
// ------------------------------------------------------------------------------------------
constexpr int ATT4_WARPS = 16;
constexpr int ATT4_ROWS = ATT4_WARPS / 2;
constexpr int ATT4_THREADS = ATT4_WARPS * 32;
constexpr int ATT4_STAGES = 3;
constexpr int ATT4_STAGE_BYTES = 4096;
constexpr int ATT4_SMEM = ATT4_WARPS * ATT4_STAGES * ATT4_STAGE_BYTES + ATT4_WARPS * ATT4_STAGES * 8 +
                          ATT4_WARPS * 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void pair_barrier(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// Row -> warp mapping shared by both v4 kernels.  Rows are dealt by rank: `order` lists the rows
// by decreasing region count (NULL = identity), rank i goes to CTA i % grid, pair slot
// (i / grid) % 8, so every SM gets the same mix of long and short rows.  Pair slot k is served by
// warps k and 15 - k: with the slots of one CTA sorted by length this also equalises the four
// warp schedulers (warp % 4).
struct Att4Map {
  int pair, half;
  __device__ __forceinline__ Att4Map(int warp) : pair(warp < 8 ? warp : 15 - warp), half(warp >> 3) {}
};

template <int AR>
__global__ void __launch_bounds__(ATT4_THREADS, 1)
attention_fwd4_kernel(const __nv_bfloat16* __restrict__ p_att16, const __nv_bfloat16* __restrict__ att_e16,
                      const int* __restrict__ off, int Lfix, const int* __restrict__ order,
                      const float* __restrict__ s_row0, int64_t lds, int att_h_col,
                      const float* __restrict__ w_alpha, __nv_bfloat16* __restrict__ att_res16,
                      float* att_w, int B, float* __restrict__ att_res32) {
  static_assert(AR == 512, "lane -> 2 x 8 columns mapping; one region row = 1 KB");
  constexpr int EPL = AR / 32;
  constexpr int ROWB = AR * 2;                    // bytes of one region row
  extern __shared__ __align__(128) uint8_t smem4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Att4Map map(warp);
  const int half = map.half;
  const int partner = 15 - warp;
  uint8_t* ring = smem4 + warp * (ATT4_STAGES * ATT4_STAGE_BYTES);
  pdl_launch_dependents();
  float ah[EPL], al[EPL];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) al[h * 8 + j] = __ldg(w_alpha + h * 256 + lane * 8 + j);
  pdl_wait();
  uint32_t it = 0;                                // ring items consumed so far (stage index)
  for (int rank = blockIdx.x + gridDim.x * map.pair; rank < B; rank += gridDim.x * ATT4_ROWS) {
    const int b = order ? order[rank] : rank;
    const int r0 = off ? off[b] : b * Lfix;
    const int Lb = off ? off[b + 1] - r0 : Lfix;
    const int n0 = (Lb + 1) >> 1;
    const int a0 = half ? n0 : 0;                 // this warp: regions [a0, a0 + n)
    const int n = half ? Lb - n0 : n0;
    const int npair = (n + 1) >> 1;
    const __nv_bfloat16* pg = p_att16 + int64_t(r0 + a0) * AR;
    const __nv_bfloat16* eg = att_e16 + int64_t(r0 + a0) * AR;
    // one ring item = a pair of regions: [p0 p1 e0 e1], 1 KB each; every lane copies 16-byte pieces
    // (cp.async, L1 bypass) and EVERY call commits a group, so the wait depth below is static
    auto issue = [&](int i) {
      if (i < npair) {
        uint8_t* dst = ring + ((it + i) % ATT4_STAGES) * ATT4_STAGE_BYTES;
        const int pieces = min(2, n - 2 * i) * (ROWB / 512);      // 512-byte warp pieces per tensor
        const uint8_t* ps = reinterpret_cast<const uint8_t*>(pg + int64_t(2 * i) * AR) + lane * 16;
        const uint8_t* es = reinterpret_cast<const uint8_t*>(eg + int64_t(2 * i) * AR) + lane * 16;
#pragma unroll
        for (int k = 0; k < 2 * (ROWB / 512); ++k)
          if (k < pieces) {
            cp_async16(dst + k * 512 + lane * 16, ps + k * 512);
            cp_async16(dst + 2 * ROWB + k * 512 + lane * 16, es + k * 512);
          }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < ATT4_STAGES; ++i) issue(i);
    {
      const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 x = *reinterpret_cast<const float4*>(att_h + h * 256 + lane * 8);
        const float4 y = *reinterpret_cast<const float4*>(att_h + h * 256 + lane * 8 + 4);
        ah[h * 8 + 0] = x.x; ah[h * 8 + 1] = x.y; ah[h * 8 + 2] = x.z; ah[h * 8 + 3] = x.w;
        ah[h * 8 + 4] = y.x; ah[h * 8 + 5] = y.y; ah[h * 8 + 6] = y.z; ah[h * 8 + 7] = y.w;
      }
    }
    float m = -INFINITY, sum = 0.f, acc[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) acc[j] = 0.f;
    float* wrow = att_w + r0 + a0;                // raw scores first, normalised in place below
    for (int i = 0; i < npair; ++i) {
      const int st = (it + i) % ATT4_STAGES;
      cp_async_wait<ATT4_STAGES - 1>();           // this lane's pieces of item i have landed
      __syncwarp();                               // ... and so have the other lanes'
      const bool two = 2 * i + 1 < n;
      const uint8_t* ps = ring + st * ATT4_STAGE_BYTES;
      const uint8_t* es = ps + 2 * ROWB;
      const int o1 = two ? ROWB : 0;              // a single-region stage re-reads region 0 (weight 0)
      float sc0 = 0.f, sc1 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(ps + (h * 256 + lane * 8) * 2);
        const uint4 u1 = *reinterpret_cast<const uint4*>(ps + o1 + (h * 256 + lane * 8) * 2);
        float f0[8], f1[8];
        bf16x8_to_float(u0, f0);
        bf16x8_to_float(u1, f1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sc0 += al[h * 8 + j] * tanh_fast(f0[j] + ah[h * 8 + j]);
          sc1 += al[h * 8 + j] * tanh_fast(f1[j] + ah[h * 8 + j]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sc0 += __shfl_xor_sync(0xffffffffu, sc0, o);
        sc1 += __shfl_xor_sync(0xffffffffu, sc1, o);
      }
      if (lane == 0) {
        wrow[2 * i] = sc0;
        if (two) wrow[2 * i + 1] = sc1;
      }
      const float nm = fmaxf(m, two ? fmaxf(sc0, sc1) : sc0);
      const float scale = __expf(m - nm);         // exp(-inf) = 0 on the first pair
      const float w0 = __expf(sc0 - nm), w1 = two ? __expf(sc1 - nm) : 0.f;
      m = nm;
      sum = sum * scale + w0 + w1;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(es + (h * 256 + lane * 8) * 2);
        const uint4 u1 = *reinterpret_cast<const uint4*>(es + o1 + (h * 256 + lane * 8) * 2);
        float f0[8], f1[8];
        bf16x8_to_float(u0, f0);
        bf16x8_to_float(u1, f1);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[h * 8 + j] = acc[h * 8 + j] * scale + w0 * f0[j] + w1 * f1[j];
      }
      __syncwarp();                               // every lane is done with this stage
      issue(i + ATT4_STAGES);
    }
    it += npair;
    cp_async_wait<0>();
    // the ring is drained: its first 2 KB + 8 B carry this warp's partial state to the partner
    float* part = reinterpret_cast<float*>(ring);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* d = part + h * 256 + lane * 8;
      *reinterpret_cast<float4*>(d) = make_float4(acc[h * 8], acc[h * 8 + 1], acc[h * 8 + 2], acc[h * 8 + 3]);
      *reinterpret_cast<float4*>(d + 4) =
          make_float4(acc[h * 8 + 4], acc[h * 8 + 5], acc[h * 8 + 6], acc[h * 8 + 7]);
    }
    if (lane == 0) { part[AR] = m; part[AR + 1] = sum; }
    pair_barrier(map.pair);
    const float* other = reinterpret_cast<const float*>(smem4 + partner * (ATT4_STAGES * ATT4_STAGE_BYTES));
    const float m2 = other[AR], s2 = other[AR + 1];
    const float M = fmaxf(m, m2);                 // the first half always holds >= 1 region
    const float f1 = __expf(m - M), f2 = __expf(m2 - M);
    const float inv = 1.f / (sum * f1 + s2 * f2);
    {
      const int c0 = half * 256 + lane * 8;       // this warp writes 256 of the 512 columns
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f1 * part[c0 + j] + f2 * other[c0 + j]) * inv;
      *reinterpret_cast<uint4*>(att_res16 + int64_t(b) * AR + c0) = float8_to_bf16x8(o);
      if (att_res32) {      // fp32 copy for the single-pass backward (softmax-backward mean, see bwd5)
        float* d = att_res32 + int64_t(b) * AR + c0;
        *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    for (int l = lane; l < n; l += 32) wrow[l] = __expf(wrow[l] - M) * inv;
    // the partner must be done with this warp's partials before the next row's copies land on them
    if (rank + gridDim.x * ATT4_ROWS < B) pair_barrier(map.pair);
  }
}

template <int AR>
__global__ void __launch_bounds__(ATT4_THREADS, 1)
attention_bwd4_kernel(const __nv_bfloat16* __restrict__ p_att16, const __nv_bfloat16* __restrict__ att_e16,
                      const int* __restrict__ off, int Lfix, const int* __restrict__ order,
                      const float* __restrict__ s_row0, int64_t lds, int att_h_col,
                      const float* __restrict__ w_alpha, const float* __restrict__ d_att_res,
                      const float* __restrict__ att_w, float* de_out, __nv_bfloat16* __restrict__ dscat,
                      int B) {
  static_assert(AR == 512, "lane -> 2 x 8 columns mapping; one region row = 1 KB");
  constexpr int EPL = AR / 32;
  constexpr int ROWB = AR * 2;
  extern __shared__ __align__(128) uint8_t smem4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Att4Map map(warp);
  const int half = map.half;
  const int partner = 15 - warp;
  uint8_t* ring = smem4 + warp * (ATT4_STAGES * ATT4_STAGE_BYTES);
  float* s_dot = reinterpret_cast<float*>(smem4 + ATT4_WARPS * ATT4_STAGES * ATT4_STAGE_BYTES +
                                          ATT4_WARPS * ATT4_STAGES * 8);
  pdl_launch_dependents();
  pdl_wait();
  uint32_t it = 0;
  for (int rank = blockIdx.x + gridDim.x * map.pair; rank < B; rank += gridDim.x * ATT4_ROWS) {
    const int b = order ? order[rank] : rank;
    const int r0 = off ? off[b] : b * Lfix;
    const int Lb = off ? off[b + 1] - r0 : Lfix;
    const int n0 = (Lb + 1) >> 1;
    const int a0 = half ? n0 : 0;
    const int n = half ? Lb - n0 : n0;
    const int ng = (n + 3) >> 2;                  // groups of 4 regions; item i < ng: att_e, else p_att
    const __nv_bfloat16* pg = p_att16 + int64_t(r0 + a0) * AR;
    const __nv_bfloat16* eg = att_e16 + int64_t(r0 + a0) * AR;
    // one ring item = 4 regions of one tensor (4 KB); every call commits a group (static wait depth)
    auto issue = [&](int i) {
      if (i < 2 * ng) {
        const int c = i < ng ? i : i - ng;
        uint8_t* dst = ring + ((it + i) % ATT4_STAGES) * ATT4_STAGE_BYTES;
        const int pieces = min(4, n - 4 * c) * (ROWB / 512);
        const uint8_t* src =
            reinterpret_cast<const uint8_t*>((i < ng ? eg : pg) + int64_t(4 * c) * AR) + lane * 16;
#pragma unroll
        for (int k = 0; k < 4 * (ROWB / 512); ++k)
          if (k < pieces) cp_async16(dst + k * 512 + lane * 16, src + k * 512);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < ATT4_STAGES; ++i) issue(i);
    const float* wrow = att_w + r0 + a0;
    float* drow = de_out + r0 + a0;               // dw first, de in place below
    // pass 1: dw_l = <d_att_res, att_e_l>, dot = sum_l w_l dw_l
    float dot = 0.f;
    {
      float dr[EPL];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float* src = d_att_res + int64_t(b) * AR + h * 256 + lane * 8;
        const float4 x = *reinterpret_cast<const float4*>(src);
        const float4 y = *reinterpret_cast<const float4*>(src + 4);
        dr[h * 8 + 0] = x.x; dr[h * 8 + 1] = x.y; dr[h * 8 + 2] = x.z; dr[h * 8 + 3] = x.w;
        dr[h * 8 + 4] = y.x; dr[h * 8 + 5] = y.y; dr[h * 8 + 6] = y.z; dr[h * 8 + 7] = y.w;
      }
      for (int c = 0; c < ng; ++c) {
        const int st = (it + c) % ATT4_STAGES;
        cp_async_wait<ATT4_STAGES - 1>();
        __syncwarp();
        const uint8_t* es = ring + st * ATT4_STAGE_BYTES;
        const int rows = min(4, n - 4 * c);
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < rows) {                          // warp-uniform
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float f[8];
              bf16x8_to_float(*reinterpret_cast<const uint4*>(es + q * ROWB + (h * 256 + lane * 8) * 2), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) a[q] += dr[h * 8 + j] * f[j];
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int q = 0; q < 4; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < rows) dot += wrow[4 * c + q] * a[q];
        const float mine = lane == 0 ? a[0] : lane == 1 ? a[1] : lane == 2 ? a[2] : a[3];
        if (lane < rows) drow[4 * c + lane] = mine;
        __syncwarp();
        issue(c + ATT4_STAGES);
      }
    }
    if (lane == 0) s_dot[warp] = dot;
    pair_barrier(map.pair);
    const float dotw = dot + s_dot[partner];
    // pass 2: de_l = w_l (dw_l - dotw);  d_att_h[j] = alpha_j sum_l de_l (1 - tanh^2(p_att[l,j] + att_h[j]))
    float ah[EPL], acc[EPL];
    {
      const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 x = *reinterpret_cast<const float4*>(att_h + h * 256 + lane * 8);
        const float4 y = *reinterpret_cast<const float4*>(att_h + h * 256 + lane * 8 + 4);
        ah[h * 8 + 0] = x.x; ah[h * 8 + 1] = x.y; ah[h * 8 + 2] = x.z; ah[h * 8 + 3] = x.w;
        ah[h * 8 + 4] = y.x; ah[h * 8 + 5] = y.y; ah[h * 8 + 6] = y.z; ah[h * 8 + 7] = y.w;
      }
    }
#pragma unroll
    for (int j = 0; j < EPL; ++j) acc[j] = 0.f;
    for (int c = 0; c < ng; ++c) {
      const int st = (it + ng + c) % ATT4_STAGES;
      cp_async_wait<ATT4_STAGES - 1>();
      __syncwarp();
      const uint8_t* ps = ring + st * ATT4_STAGE_BYTES;
      const int rows = min(4, n - 4 * c);
      float de[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) de[q] = (q < rows) ? wrow[4 * c + q] * (drow[4 * c + q] - dotw) : 0.f;
      __syncwarp();                               // all lanes have read dw before it becomes de
      {
        const float mine = lane == 0 ? de[0] : lane == 1 ? de[1] : lane == 2 ? de[2] : de[3];
        if (lane < rows) drow[4 * c + lane] = mine;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < rows) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float f[8];
            bf16x8_to_float(*reinterpret_cast<const uint4*>(ps + q * ROWB + (h * 256 + lane * 8) * 2), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t = tanh_fast(f[j] + ah[h * 8 + j]);
              acc[h * 8 + j] += de[q] * (1.f - t * t);
            }
          }
        }
      }
      __syncwarp();
      issue(ng + c + ATT4_STAGES);
    }
    it += 2 * ng;
    cp_async_wait<0>();
    float* part = reinterpret_cast<float*>(ring);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* d = part + h * 256 + lane * 8;
      *reinterpret_cast<float4*>(d) = make_float4(acc[h * 8], acc[h * 8 + 1], acc[h * 8 + 2], acc[h * 8 + 3]);
      *reinterpret_cast<float4*>(d + 4) =
          make_float4(acc[h * 8 + 4], acc[h * 8 + 5], acc[h * 8 + 6], acc[h * 8 + 7]);
    }
    pair_barrier(map.pair);
    {
      const float* other = reinterpret_cast<const float*>(smem4 + partner * (ATT4_STAGES * ATT4_STAGE_BYTES));
      const int c0 = half * 256 + lane * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (part[c0 + j] + other[c0 + j]) * __ldg(w_alpha + c0 + j);
      *reinterpret_cast<uint4*>(dscat + int64_t(b) * lds + att_h_col + c0) = float8_to_bf16x8(o);
    }
    if (rank + gridDim.x * ATT4_ROWS < B) pair_barrier(map.pair);
  }
}

// ------------------------------------------------------------------------------------------
// v5 backward: ONE pass.  The softmax backward needs dot = sum_l w_l dw_l with dw_l = <d_att_res,
// att_e_l>, which is why v4 walks att_e first and p_att second.  But
//     sum_l w_l <d_att_res, att_e_l> = <d_att_res, sum_l w_l att_e_l> = <d_att_res, att_res>,
// and att_res is what the forward pass produced (kept in fp32 for this purpose): the mean is known
// before the first region arrives, so (att_e, p_att) stream through together exactly like in the
// forward kernel -- ring items [p0 p1 e0 e1], no barrier between the two halves of a row until
// their d_att_h partials meet.  Same arithmetic otherwise:
//     de_l = w_l (dw_l - dot) ;  d_att_h[j] = alpha_j sum_l de_l (1 - tanh^2(p_att[l,j] + att_h[j]))
// ------------------------------------------------------------------------------------------
template <int AR>
__global__ void __launch_bounds__(ATT4_THREADS, 1)
attention_bwd5_kernel(const __nv_bfloat16* __restrict__ p_att16, const __nv_bfloat16* __restrict__ att_e16,
                      const int* __restrict__ off, int Lfix, const int* __restrict__ order,
                      const float* __restrict__ s_row0, int64_t lds, int att_h_col,
                      const float* __restrict__ w_alpha, const float* __restrict__ d_att_res,
                      const float* __restrict__ att_res32, const float* __restrict__ att_w,
                      float* de_out, __nv_bfloat16* __restrict__ dscat, int B) {
  static_assert(AR == 512, "lane -> 2 x 8 columns mapping; one region row = 1 KB");
  constexpr int EPL = AR / 32;
  constexpr int ROWB = AR * 2;
  extern __shared__ __align__(128) uint8_t smem4[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Att4Map map(warp);
  const int half = map.half;
  const int partner = 15 - warp;
  uint8_t* ring = smem4 + warp * (ATT4_STAGES * ATT4_STAGE_BYTES);
  pdl_launch_dependents();
  pdl_wait();
  uint32_t it = 0;
  for (int rank = blockIdx.x + gridDim.x * map.pair; rank < B; rank += gridDim.x * ATT4_ROWS) {
    const int b = order ? order[rank] : rank;
    const int r0 = off ? off[b] : b * Lfix;
    const int Lb = off ? off[b + 1] - r0 : Lfix;
    const int n0 = (Lb + 1) >> 1;
    const int a0 = half ? n0 : 0;                 // this warp: regions [a0, a0 + n)
    const int n = half ? Lb - n0 : n0;
    const int npair = (n + 1) >> 1;
    const __nv_bfloat16* pg = p_att16 + int64_t(r0 + a0) * AR;
    const __nv_bfloat16* eg = att_e16 + int64_t(r0 + a0) * AR;
    auto issue = [&](int i) {
      if (i < npair) {
        uint8_t* dst = ring + ((it + i) % ATT4_STAGES) * ATT4_STAGE_BYTES;
        const int pieces = min(2, n - 2 * i) * (ROWB / 512);
        const uint8_t* ps = reinterpret_cast<const uint8_t*>(pg + int64_t(2 * i) * AR) + lane * 16;
        const uint8_t* es = reinterpret_cast<const uint8_t*>(eg + int64_t(2 * i) * AR) + lane * 16;
#pragma unroll
        for (int k = 0; k < 2 * (ROWB / 512); ++k)
          if (k < pieces) {
            cp_async16(dst + k * 512 + lane * 16, ps + k * 512);
            cp_async16(dst + 2 * ROWB + k * 512 + lane * 16, es + k * 512);
          }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < ATT4_STAGES; ++i) issue(i);
    // per-row vectors: d_att_res, att_h; dot = <d_att_res, att_res> (every warp covers all columns)
    float dr[EPL], ah[EPL], acc[EPL];
    float dot = 0.f;
    {
      const float* att_h = s_row0 + int64_t(b) * lds + att_h_col;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = h * 256 + lane * 8;
        const float4 x = *reinterpret_cast<const float4*>(d_att_res + int64_t(b) * AR + c0);
        const float4 y = *reinterpret_cast<const float4*>(d_att_res + int64_t(b) * AR + c0 + 4);
        const float4 rx = *reinterpret_cast<const float4*>(att_res32 + int64_t(b) * AR + c0);
        const float4 ry = *reinterpret_cast<const float4*>(att_res32 + int64_t(b) * AR + c0 + 4);
        const float4 hx = *reinterpret_cast<const float4*>(att_h + c0);
        const float4 hy = *reinterpret_cast<const float4*>(att_h + c0 + 4);
        dr[h * 8 + 0] = x.x; dr[h * 8 + 1] = x.y; dr[h * 8 + 2] = x.z; dr[h * 8 + 3] = x.w;
        dr[h * 8 + 4] = y.x; dr[h * 8 + 5] = y.y; dr[h * 8 + 6] = y.z; dr[h * 8 + 7] = y.w;
        ah[h * 8 + 0] = hx.x; ah[h * 8 + 1] = hx.y; ah[h * 8 + 2] = hx.z; ah[h * 8 + 3] = hx.w;
        ah[h * 8 + 4] = hy.x; ah[h * 8 + 5] = hy.y; ah[h * 8 + 6] = hy.z; ah[h * 8 + 7] = hy.w;
        dot += x.x * rx.x + x.y * rx.y + x.z * rx.z + x.w * rx.w + y.x * ry.x + y.y * ry.y + y.z * ry.z +
               y.w * ry.w;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
    for (int j = 0; j < EPL; ++j) acc[j] = 0.f;
    const float* wrow = att_w + r0 + a0;
    float* drow = de_out + r0 + a0;
    for (int i = 0; i < npair; ++i) {
      const int st = (it + i) % ATT4_STAGES;
      cp_async_wait<ATT4_STAGES - 1>();
      __syncwarp();
      const bool two = 2 * i + 1 < n;
      const uint8_t* ps = ring + st * ATT4_STAGE_BYTES;
      const uint8_t* es = ps + 2 * ROWB;
      const int o1 = two ? ROWB : 0;              // a single-region stage re-reads region 0 (de = 0)
      const float w0 = wrow[2 * i], w1 = two ? wrow[2 * i + 1] : 0.f;
      float dw0 = 0.f, dw1 = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f0[8], f1[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(es + (h * 256 + lane * 8) * 2), f0);
        bf16x8_to_float(*reinterpret_cast<const uint4*>(es + o1 + (h * 256 + lane * 8) * 2), f1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dw0 += dr[h * 8 + j] * f0[j];
          dw1 += dr[h * 8 + j] * f1[j];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dw0 += __shfl_xor_sync(0xffffffffu, dw0, o);
        dw1 += __shfl_xor_sync(0xffffffffu, dw1, o);
      }
      const float de0 = w0 * (dw0 - dot), de1 = two ? w1 * (dw1 - dot) : 0.f;
      if (lane == 0) {
        drow[2 * i] = de0;
        if (two) drow[2 * i + 1] = de1;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f0[8], f1[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(ps + (h * 256 + lane * 8) * 2), f0);
        bf16x8_to_float(*reinterpret_cast<const uint4*>(ps + o1 + (h * 256 + lane * 8) * 2), f1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t0 = tanh_fast(f0[j] + ah[h * 8 + j]);
          const float t1 = tanh_fast(f1[j] + ah[h * 8 + j]);
          acc[h * 8 + j] += de0 * (1.f - t0 * t0) + de1 * (1.f - t1 * t1);
        }
      }
      __syncwarp();                               // every lane is done with this stage
      issue(i + ATT4_STAGES);
    }
    it += npair;
    cp_async_wait<0>();
    float* part = reinterpret_cast<float*>(ring);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* d = part + h * 256 + lane * 8;
      *reinterpret_cast<float4*>(d) = make_float4(acc[h * 8], acc[h * 8 + 1], acc[h * 8 + 2], acc[h * 8 + 3]);
      *reinterpret_cast<float4*>(d + 4) =
          make_float4(acc[h * 8 + 4], acc[h * 8 + 5], acc[h * 8 + 6], acc[h * 8 + 7]);
    }
    pair_barrier(map.pair);
    {
      const float* other = reinterpret_cast<const float*>(smem4 + partner * (ATT4_STAGES * ATT4_STAGE_BYTES));
      const int c0 = half * 256 + lane * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (part[c0 + j] + other[c0 + j]) * __ldg(w_alpha + c0 + j);
      *reinterpret_cast<uint4*>(dscat + int64_t(b) * lds + att_h_col + c0) = float8_to_bf16x8(o);
    }
    if (rank + gridDim.x * ATT4_ROWS < B) pair_barrier(map.pair);
  }
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
                           float* __restrict__ galpha_part, float* __restrict__ gbias_part,
                           const int* __restrict__ order) {
  static_assert(AR == 2 * ATT_THREADS, "thread -> 2 columns mapping");
  constexpr int CHUNK_BYTES = ATT_CH * AR * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  // CTAs are dispatched in blockIdx order: longest rows first keeps the tail of the grid short
  const int b = order ? order[blockIdx.x] : blockIdx.x;
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
