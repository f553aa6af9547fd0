// Library runtime: last-error string, device query, TMA tensor-map encoding (driver entry point
// resolved through the runtime so the library does not link libcuda directly).
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <utility>
#include <vector>

#include "../../include/coopcap.h"
#include <algorithm>
#include "common.cuh"
#include "gemm.cuh"

namespace coopcap {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

// ---- launch counter + event timeline -----------------------------------------------------------
struct ProfRec { int kind; cudaEvent_t ev; double flops, bytes; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t g_prof_begin = nullptr;
static bool g_prof_on = false;
static long long g_launches[PROF_NKINDS] = {0};

void prof_mark(int kind, cudaStream_t stream, double flops, double bytes) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (kind >= 0 && kind < PROF_NKINDS) ++g_launches[kind];
  if (!g_prof_on || stream == CC_NO_STREAM) return;
  cudaEvent_t ev;
  if (!g_prof_pool.empty()) { ev = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, stream);
  g_prof.push_back({kind, ev, flops, bytes});
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("COOPCAP_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int ensure_dyn_smem(const void* func, int bytes) {
  if (bytes <= 48 * 1024) return CC_OK;           // within the default limit: nothing to opt in to
  static std::mutex mu;
  static std::vector<std::pair<std::pair<int, const void*>, int>> seen;   // ((device, func), bytes)
  int dev = 0;
  CC_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  for (auto& e : seen)
    if (e.first.first == dev && e.first.second == func) {
      if (e.second >= bytes) return CC_OK;
      CC_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      e.second = bytes;
      return CC_OK;
    }
  CC_CHECK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  seen.push_back({{dev, func}, bytes});
  return CC_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// Encoded tensor maps are pure functions of (base, element size, extents, pitch, box): the training
// loop re-encodes the same few hundred descriptors every step (the caching allocator hands the same
// blocks back), and cuTensorMapEncodeTiled was a visible share of the ~17 us of host time per
// launch.  Direct-mapped, per-thread cache: no locks, a collision just re-encodes.
struct TmapKey {
  const void* base; int64_t rows, cols, ld; int elem_bytes, box_rows, box_cols, valid;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld &&
           elem_bytes == o.elem_bytes && box_rows == o.box_rows && box_cols == o.box_cols && valid == o.valid;
  }
};
struct TmapSlot { TmapKey key; CUtensorMap map; };
constexpr int TMAP_SLOTS = 2048;
static int encode_tmap_2d_uncached(CUtensorMap* out, const void* base, int elem_bytes, int64_t rows,
                                   int64_t cols, int64_t ld, int box_rows, int box_cols);

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, int64_t rows, int64_t cols,
                   int64_t ld, int box_rows, int box_cols) {
  static thread_local std::vector<TmapSlot> cache(TMAP_SLOTS);
  const TmapKey key = {base, rows, cols, ld, elem_bytes, box_rows, box_cols, 1};
  uint64_t h = reinterpret_cast<uintptr_t>(base) >> 4;
  h = (h ^ uint64_t(rows) * 0x9E3779B97F4A7C15ull ^ uint64_t(cols) * 0xC2B2AE3D27D4EB4Full ^
       uint64_t(ld) * 0x165667B19E3779F9ull ^ uint64_t(box_rows * 131 + box_cols * 7 + elem_bytes)) *
      0xD6E8FEB86659FD93ull;
  TmapSlot& slot = cache[(h >> 40) % TMAP_SLOTS];
  if (slot.key == key) {
    *out = slot.map;
    return CC_OK;
  }
  const int rc = encode_tmap_2d_uncached(out, base, elem_bytes, rows, cols, ld, box_rows, box_cols);
  if (rc == CC_OK) { slot.key = key; slot.map = *out; }
  return rc;
}

static int encode_tmap_2d_uncached(CUtensorMap* out, const void* base, int elem_bytes, int64_t rows,
                                   int64_t cols, int64_t ld, int box_rows, int box_cols) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled driver entry point not available");
    return CC_ERR_DRIVER;
  }
  CC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  CC_REQUIRE((ld * elem_bytes) % 16 == 0, "TMA row pitch %lld B not a multiple of 16",
             (long long)(ld * elem_bytes));
  CC_REQUIRE(box_cols * elem_bytes == 128 && box_rows <= 256 && box_rows > 0,
             "TMA box %dx%d unsupported", box_rows, box_cols);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt =
      elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d",
                   (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    return CC_ERR_DRIVER;
  }
  return CC_OK;
}

}  // namespace coopcap

extern "C" {

const char* coopcap_last_error(void) { return coopcap::g_last_error; }

int coopcap_version(void) { return COOPCAP_VERSION; }

int coopcap_h2d_ragged_rows(void* dst, const void* src_host, const int* lens_host, int B,
                            int64_t row_stride_bytes, int64_t unit_bytes, coopcap_stream_t stream) {
  using namespace coopcap;
  CC_REQUIRE(dst && src_host && lens_host && B >= 0, "h2d_ragged_rows: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // coalesce runs of full rows into one copy; otherwise one copy per row
  int b = 0;
  while (b < B) {
    int64_t bytes = int64_t(lens_host[b]) * unit_bytes;
    CC_REQUIRE(bytes >= 0 && bytes <= row_stride_bytes, "h2d_ragged_rows: row %d has %lld B > pitch",
               b, (long long)bytes);
    int e = b + 1;
    if (bytes == row_stride_bytes) {
      while (e < B && int64_t(lens_host[e]) * unit_bytes == row_stride_bytes) ++e;
      bytes = int64_t(e - b) * row_stride_bytes;
    }
    if (bytes > 0)
      CC_CHECK_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + int64_t(b) * row_stride_bytes,
                                    static_cast<const char*>(src_host) + int64_t(b) * row_stride_bytes,
                                    size_t(bytes), cudaMemcpyHostToDevice, s));
    b = e;
  }
  return CC_OK;
}

long long coopcap_launch_count(void) {
  std::lock_guard<std::mutex> lk(coopcap::g_prof_mu);
  long long t = 0;
  for (int i = 0; i < coopcap::PROF_NKINDS; ++i) t += coopcap::g_launches[i];
  return t;
}

int coopcap_prof_kinds(void) { return coopcap::PROF_NKINDS; }

int coopcap_prof_enable(int on, coopcap_stream_t stream) {
  using namespace coopcap;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) g_prof_pool.push_back(r.ev);
  g_prof.clear();
  g_prof_on = on != 0;
  if (g_prof_on) {
    if (!g_prof_begin) CC_CHECK_CUDA(cudaEventCreate(&g_prof_begin));
    CC_CHECK_CUDA(cudaEventRecord(g_prof_begin, reinterpret_cast<cudaStream_t>(stream)));
  }
  return CC_OK;
}

int coopcap_prof_report(double* ms, double* flops, double* bytes, long long* launches, int nkinds) {
  using namespace coopcap;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  CC_REQUIRE(nkinds >= PROF_NKINDS, "prof_report: need %d slots", PROF_NKINDS);
  for (int i = 0; i < nkinds; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
  if (g_prof.empty()) return CC_OK;
  CC_CHECK_CUDA(cudaEventSynchronize(g_prof.back().ev));
  cudaEvent_t prev = g_prof_begin;
  for (auto& r : g_prof) {
    float t = 0.f;
    CC_CHECK_CUDA(cudaEventElapsedTime(&t, prev, r.ev));
    ms[r.kind] += t; flops[r.kind] += r.flops; bytes[r.kind] += r.bytes; ++launches[r.kind];
    prev = r.ev;
  }
  for (auto& r : g_prof) g_prof_pool.push_back(r.ev);
  g_prof.clear();
  return CC_OK;
}

int coopcap_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(coopcap_gemm_args);
    case 1: return (int)sizeof(coopcap_speaker_pack);
    case 2: return (int)sizeof(coopcap_speaker);
    case 3: return (int)sizeof(coopcap_speaker_grads);
    case 4: return (int)sizeof(coopcap_listener_pack);
    case 5: return (int)sizeof(coopcap_listener);
    case 6: return (int)sizeof(coopcap_listener_grads);
    case 7: return (int)sizeof(coopcap_cider);
    case 8: return (int)sizeof(coopcap_beam);
    default: return -1;
  }
}

int coopcap_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  CC_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CC_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return coopcap::CC_OK;
}

int coopcap_l2_persist(const void* base, int64_t bytes, int64_t* granted, coopcap_stream_t stream) {
  using namespace coopcap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, max_persist = 0, max_window = 0;
  CC_CHECK_CUDA(cudaGetDevice(&dev));
  CC_CHECK_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  CC_CHECK_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  cudaStreamAttrValue v = {};
  if (bytes <= 0 || base == nullptr) {
    v.accessPolicyWindow.num_bytes = 0;
    v.accessPolicyWindow.hitRatio = 0.f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
    v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    CC_CHECK_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
    CC_CHECK_CUDA(cudaCtxResetPersistingL2Cache());
    CC_CHECK_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
    if (granted) *granted = 0;
    return CC_OK;
  }
  const int64_t aside = std::min<int64_t>(bytes, max_persist);
  const int64_t window = std::min<int64_t>(bytes, max_window);
  if (aside <= 0 || window <= 0) {           // the device has no persisting L2: nothing to do
    if (granted) *granted = 0;
    return CC_OK;
  }
  CC_CHECK_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(aside)));
  v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
  v.accessPolicyWindow.num_bytes = size_t(window);
  v.accessPolicyWindow.hitRatio = float(std::min(1.0, double(aside) / double(window)));
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  CC_CHECK_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
  if (granted) *granted = aside;
  return CC_OK;
}

}  // extern "C"
