// C ABI for the dense contraction engine: dispatch over operand kind / majors / tile width, the
// SIMT cross-check kernel, and the fp32->bf16 cast.
#include "../../include/coopcap.h"
#include "common.cuh"
#include "gemm.cuh"

namespace coopcap {

// ------------------------------------------------------------------------------------------
// SIMT cross-check kernel (tests only): one thread per output element, fp32 math.
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) {
  return *p;
}

template <typename T>
__global__ void gemm_simt_kernel(const T* __restrict__ A, int64_t lda, int a_major,
                                 const T* __restrict__ B, int64_t ldb, int b_major, int M, int N,
                                 int K, EpiStoreParams p) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y * blockDim.y + threadIdx.y;
  if (m >= M || n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = a_major == 0 ? ld_as_float(A + int64_t(m) * lda + k)
                                 : ld_as_float(A + int64_t(k) * lda + m);
    const float b = b_major == 0 ? ld_as_float(B + int64_t(n) * ldb + k)
                                 : ld_as_float(B + int64_t(k) * ldb + n);
    acc = fmaf(a, b, acc);
  }
  float x = acc * p.alpha;
  if (p.bias) x += p.bias[n];
  if (p.relu) x = fmaxf(x, 0.f);
  if (p.row_scale) x *= p.row_scale[m];
  if (p.C) {
    float* d = p.C + int64_t(m) * p.ldc + n;
    if (p.mode == 0) *d = x;
    else if (p.mode == 1) *d += x;
    else atomicAdd(d, x);
  }
  if (p.C16) p.C16[int64_t(m) * p.ldc16 + n] = __float2bfloat16_rn(x);
  if (p.Ct16) p.Ct16[int64_t(n) * p.ldct + m] = __float2bfloat16_rn(x);
}

template <int KIND, int BN>
static int dispatch_major(const coopcap_gemm_args* a, const EpiStoreParams& ep, cudaStream_t s) {
  if (a->a_major == 0 && a->b_major == 0)
    return launch_gemm_tc<KIND, BN, 0, 0, EpiStore>(a->A, a->lda, a->B, a->ldb, a->M, a->N, a->K,
                                                    a->split_k, ep, s);
  if (a->a_major == 0 && a->b_major == 1)
    return launch_gemm_tc<KIND, BN, 0, 1, EpiStore>(a->A, a->lda, a->B, a->ldb, a->M, a->N, a->K,
                                                    a->split_k, ep, s);
  if (a->a_major == 1 && a->b_major == 1)
    return launch_gemm_tc<KIND, BN, 1, 1, EpiStore>(a->A, a->lda, a->B, a->ldb, a->M, a->N, a->K,
                                                    a->split_k, ep, s);
  if (a->a_major == 1 && a->b_major == 0)
    return launch_gemm_tc<KIND, BN, 1, 0, EpiStore>(a->A, a->lda, a->B, a->ldb, a->M, a->N, a->K,
                                                    a->split_k, ep, s);
  set_last_error("gemm: bad majors %d/%d", a->a_major, a->b_major);
  return CC_ERR_ARG;
}

// Tile width from a small cost model: persistent CTAs process ceil(tiles / SMs) waves of tiles; a
// tile costs max(MMA time, epilogue time) (they overlap through the double-buffered accumulator)
// plus a pipeline-fill term.  Narrow tiles fill the machine better, wide tiles feed the tensor
// pipe better (operand smem traffic per MAC falls with N).
int pick_tile_n(int M, int N, int K, int split_k, int requested) {
  if (requested == 64 || requested == 128 || requested == 192 || requested == 256) return requested;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int sms = num_sms();
  const int cand[4] = {64, 128, 192, 256};
  const double eff[4] = {0.55, 0.80, 0.92, 1.00};
  double best = 1e30;
  int best_bn = 128;
  const double kk = double((K + split_k - 1) / split_k);
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (bn > 64 && bn >= 2 * N) continue;   // do not pad tiny N into a wide tile
    const int tiles = num_m * ((N + bn - 1) / bn) * split_k;
    const int waves = (tiles + sms - 1) / sms;
    const double mma = kk * bn * 1.68e-11 / eff[i];
    const double epi = bn * 5.1e-9;
    const double cost = waves * (mma > epi ? mma : epi) + (mma > epi ? epi : mma) + 1.5e-6;
    if (cost < best) { best = cost; best_bn = bn; }
  }
  return best_bn;
}

// internal entry used by the orchestration code (speaker.cu, listener.cu)
int gemm_run(int kind, int a_major, int b_major, const void* A, int64_t lda, const void* B,
             int64_t ldb, int M, int N, int K, int split_k, int tile_n, const EpiStoreParams& ep,
             cudaStream_t s) {
  coopcap_gemm_args a = {};
  a.kind = kind; a.a_major = a_major; a.b_major = b_major;
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb;
  a.M = M; a.N = N; a.K = K; a.split_k = split_k < 1 ? 1 : split_k;
  CC_REQUIRE(kind == 0 || (a_major == 0 && b_major == 0),
             "gemm: tf32 operands must be K-major (MN-major tf32 needs the 32B-atom swizzle)");
  CC_REQUIRE(a.split_k <= 1 || (ep.mode == 2 && ep.C16 == nullptr && ep.Ct16 == nullptr),
             "gemm: split_k > 1 needs mode 2 (atomicAdd) and fp32 output only");
  const int bn = pick_tile_n(M, N, K, a.split_k, tile_n);
  if (kind == 0) {
    if (bn == 256) return dispatch_major<0, 256>(&a, ep, s);
    if (bn == 192) return dispatch_major<0, 192>(&a, ep, s);
    if (bn == 128) return dispatch_major<0, 128>(&a, ep, s);
    return dispatch_major<0, 64>(&a, ep, s);
  }
  if (bn == 256) return launch_gemm_tc<1, 256, 0, 0, EpiStore>(A, lda, B, ldb, M, N, K, a.split_k, ep, s);
  if (bn == 128) return launch_gemm_tc<1, 128, 0, 0, EpiStore>(A, lda, B, ldb, M, N, K, a.split_k, ep, s);
  return launch_gemm_tc<1, 64, 0, 0, EpiStore>(A, lda, B, ldb, M, N, K, a.split_k, ep, s);
}

int gemm_store(const coopcap_gemm_args* a, cudaStream_t s) {
  CC_REQUIRE(a != nullptr, "gemm: null args");
  CC_REQUIRE(a->kind == 0 || a->kind == 1, "gemm: kind %d", a->kind);
  EpiStoreParams ep = {};
  ep.C = a->C;
  ep.C16 = reinterpret_cast<__nv_bfloat16*>(a->C16);
  ep.Ct16 = reinterpret_cast<__nv_bfloat16*>(a->Ct16);
  ep.bias = a->bias;
  ep.row_scale = a->row_scale;
  ep.ldc = a->ldc;
  ep.ldc16 = a->ldc16;
  ep.ldct = a->ldct;
  ep.alpha = a->alpha;
  ep.relu = a->relu;
  ep.mode = a->mode;
  if (a->backend == 1) {
    dim3 blk(32, 8), grd((a->N + 31) / 32, (a->M + 7) / 8);
    if (a->kind == 0)
      gemm_simt_kernel<__nv_bfloat16><<<grd, blk, 0, s>>>(
          reinterpret_cast<const __nv_bfloat16*>(a->A), a->lda, a->a_major,
          reinterpret_cast<const __nv_bfloat16*>(a->B), a->ldb, a->b_major, a->M, a->N, a->K, ep);
    else
      gemm_simt_kernel<float><<<grd, blk, 0, s>>>(reinterpret_cast<const float*>(a->A), a->lda,
                                                  a->a_major, reinterpret_cast<const float*>(a->B),
                                                  a->ldb, a->b_major, a->M, a->N, a->K, ep);
    CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
    return CC_OK;
  }
  return gemm_run(a->kind, a->a_major, a->b_major, a->A, a->lda, a->B, a->ldb, a->M, a->N, a->K,
                  a->split_k, a->tile_n, ep, s);
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, int64_t rows, int64_t cols,
                                 int64_t ld_src, __nv_bfloat16* __restrict__ dst, int64_t ld_dst,
                                 __nv_bfloat16* __restrict__ dst_t, int64_t ld_dst_t) {
  // 32x32 tiles through shared memory so both the straight and the transposed store coalesce
  __shared__ float tile[32][33];
  const int64_t c0 = int64_t(blockIdx.x) * 32, r0 = int64_t(blockIdx.y) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = src[r * ld_src + c];
      if (dst) dst[r * ld_dst + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  if (!dst_t) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst_t[c * ld_dst_t + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

__global__ void sm_clock_kernel(float* out, long long spin_ns) {
  if (threadIdx.x != 0) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while ((long long)(t1 - t0) < spin_ns);
  const long long c1 = clock64();
  *out = float(double(c1 - c0) / double(t1 - t0) * 1e3);
}

}  // namespace coopcap

extern "C" {

int coopcap_measure_sm_clock(float* mhz_out, int spin_ns, coopcap_stream_t stream) {
  using namespace coopcap;
  CC_REQUIRE(mhz_out != nullptr, "measure_sm_clock: null output");
  if (spin_ns <= 0) spin_ns = 20000;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  sm_clock_kernel<<<1, 32, 0, s>>>(mhz_out, spin_ns);
  CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  return CC_OK;
}

int coopcap_gemm(const coopcap_gemm_args* args, coopcap_stream_t stream) {
  return coopcap::gemm_store(args, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_cast_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst,
                      int64_t ld_dst, void* dst_t, int64_t ld_dst_t, coopcap_stream_t stream) {
  using namespace coopcap;
  if (rows <= 0 || cols <= 0) return CC_OK;
  dim3 blk(32, 8), grd((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  CC_REQUIRE(grd.y <= 65535, "cast_bf16: too many rows for one launch (%lld)", (long long)rows);
  cast_bf16_kernel<<<grd, blk, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, rows, cols, ld_src, reinterpret_cast<__nv_bfloat16*>(dst), ld_dst,
      reinterpret_cast<__nv_bfloat16*>(dst_t), ld_dst_t);
  CC_LAUNCH_CHECK_K(PROF_PACK, reinterpret_cast<cudaStream_t>(stream), 0.0, 0.0);
  return CC_OK;
}

}  // extern "C"
