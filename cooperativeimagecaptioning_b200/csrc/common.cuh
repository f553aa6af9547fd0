// Shared device/host helpers for libcoopcap (sm_100a only).
// PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), Philox.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace coopcap {

// ---------------------------------------------------------------------------------------------
// error plumbing: no exceptions cross the C ABI; every entry point returns 0 or a negative code.
// ---------------------------------------------------------------------------------------------
enum : int {
  CC_OK = 0,
  CC_ERR_CUDA = -1,
  CC_ERR_ARG = -2,
  CC_ERR_DRIVER = -3,
  CC_ERR_UNSUPPORTED = -4,
};

void set_last_error(const char* fmt, ...);

#define CC_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::coopcap::set_last_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,               \
                                cudaGetErrorName(_e), cudaGetErrorString(_e));               \
      return ::coopcap::CC_ERR_CUDA;                                                         \
    }                                                                                        \
  } while (0)

#define CC_REQUIRE(cond, ...)                                                                \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::coopcap::set_last_error(__VA_ARGS__);                                                \
      return ::coopcap::CC_ERR_ARG;                                                          \
    }                                                                                        \
  } while (0)

// Kernel classes for the launch counter / event timeline (coopcap_prof_* in include/coopcap.h).
enum ProfKind : int {
  PROF_MISC = 0,
  PROF_GEMM,
  PROF_ATT_FWD,
  PROF_ATT_BWD,
  PROF_ATT_DEFERRED,
  PROF_LSTM,
  PROF_SAMPLE,
  PROF_ST_BWD,
  PROF_LOGP_BWD,
  PROF_GRU,
  PROF_HINGE,
  PROF_REDUCE,
  PROF_PACK,
  PROF_ADAM,
  PROF_LOGIT_SAMPLE,   // the logit GEMM with the sampler in its epilogue: a tcgen05 contraction bound by ALU issue
  PROF_NKINDS
};
// counts the launch; when the timeline is enabled also records an event on `stream` so that the
// time between consecutive markers (one stream, back-to-back launches) is this launch's duration
void prof_mark(int kind, cudaStream_t stream, double flops, double bytes);

#define CC_LAUNCH_CHECK_K(kind, stream, flops, bytes)                                        \
  do {                                                                                       \
    CC_CHECK_CUDA(cudaGetLastError());                                                       \
    ::coopcap::prof_mark((kind), (stream), (flops), (bytes));                                \
  } while (0)
// torch's default stream is the null handle, so "no stream known" needs its own sentinel
#define CC_NO_STREAM (reinterpret_cast<cudaStream_t>(static_cast<intptr_t>(-1)))
#define CC_LAUNCH_CHECK() CC_LAUNCH_CHECK_K(::coopcap::PROF_MISC, CC_NO_STREAM, 0.0, 0.0)

int num_sms();

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  The attribute is per
// (device, function); the cache is keyed the same way, so a process that drives several GPUs (or
// switches devices between calls) sets it on each of them.  Returns CC_OK or a negative code.
int ensure_dyn_smem(const void* func, int bytes);

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL): kernels of the per-step loops are launched with the
// programmatic-stream-serialization attribute, so kernel N+1 may be scheduled while kernel N is
// still running.  Every such kernel (1) lets its own dependents go early and (2) executes
// griddepcontrol.wait -- which returns only when the preceding grid has completed and its memory
// is visible -- before it reads or writes any global data.  Both are no-ops for normal launches.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // runtime.cu: false when the environment sets COOPCAP_NO_PDL=1

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;   // COOPCAP_NO_PDL=1 falls back to plain stream order
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tile load: coordinates are (c0 = innermost/contiguous dim, c1 = outer dim), in elements.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2D tile store / reduce-add from shared memory (bulk async-group completion).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0,
                                             int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src,
                                                  int32_t c0, int32_t c1) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrives (count 1) on the mbarrier once all previously issued MMAs have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers f16/bf16 inputs, kind::tf32 fp32-as-tf32.
template <int KIND>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                        uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == 0) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// TMEM -> registers: this thread's lane (row), 32 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// same, 16 columns (half the registers: used where the epilogue math needs the rest)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, 128-byte swizzle (layout_type 2), sm_100 version bit.
//   K-major : 8-row groups of 128 B rows; SBO = 1024 B between groups; LBO unused (encoded 1).
//   MN-major: atoms of (128 B along MN) x (8 K-rows); LBO = bytes between MN atoms, SBO = 1024 B
//             between 8-K-row groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor (upper 32 bits of the 64-bit idesc operand are what the PTX takes)
template <int KIND>
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t fmt = (KIND == 0) ? 1u /*BF16*/ : 2u /*TF32*/;
  return (1u << 4) /*D = F32*/ | (fmt << 7) | (fmt << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG; backward passes regenerate the same stream)
// ---------------------------------------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  // ROUNDS = 10 is the conservative default (cuRAND's); 7 is the smallest round count of
  // Philox4x32 that passes BigCrush (Salmon et al., SC'11, table 2) and is what the per-logit noise
  // site uses: there the generator is ~1/4 of the sampling epilogue's instructions.
  template <int ROUNDS>
  __host__ __device__ static inline uint4 gen_r(uint64_t seed, uint64_t stream, uint64_t ctr) {
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32);
    uint32_t c2 = static_cast<uint32_t>(stream), c3 = static_cast<uint32_t>(stream >> 32);
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      // one 32x32->64 multiply per product (IMAD.WIDE), not a hi/lo pair
      const uint64_t p0 = uint64_t(M0) * c0, p1 = uint64_t(M1) * c2;
      const uint32_t hi0 = uint32_t(p0 >> 32), lo0 = uint32_t(p0);
      const uint32_t hi1 = uint32_t(p1 >> 32), lo1 = uint32_t(p1);
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  __host__ __device__ static inline uint4 gen(uint64_t seed, uint64_t stream, uint64_t ctr) {
    return gen_r<10>(seed, stream, ctr);
  }
  // uniform on the open interval: (k + 1/2) * 2^-23, k in [0, 2^23), i.e. [2^-24, 1 - 2^-24], every
  // value exact in fp32.  Neither end is reachable, so -log(u) and -log(1 - u) (the Gumbel and
  // exponential-race transforms) are finite and non-zero for every draw: with a closed lower end
  // u == 0 made the race score +inf once per 2^24 logits and that token won whatever its logit.
  __host__ __device__ static inline float u01(uint32_t x) {
#ifdef __CUDA_ARCH__
    // mantissa trick: [1, 2) from the top 23 bits, then shift to (0, 1): 2 instructions, same value
    return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994f;     // 1 - 2^-24
#else
    return (float(x >> 9) + 0.5f) * (1.0f / 8388608.0f);
#endif
  }
};

}  // namespace coopcap
