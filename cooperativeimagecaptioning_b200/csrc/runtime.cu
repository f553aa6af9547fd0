// Library runtime: last-error string, device query, TMA tensor-map encoding (driver entry point
// resolved through the runtime so the library does not link libcuda directly).
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/coopcap.h"
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

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, int64_t rows, int64_t cols,
                   int64_t ld, int box_rows, int box_cols) {
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

int coopcap_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(coopcap_gemm_args);
    case 1: return (int)sizeof(coopcap_speaker_pack);
    case 2: return (int)sizeof(coopcap_speaker);
    case 3: return (int)sizeof(coopcap_speaker_grads);
    case 4: return (int)sizeof(coopcap_listener_pack);
    case 5: return (int)sizeof(coopcap_listener);
    case 6: return (int)sizeof(coopcap_listener_grads);
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

}  // extern "C"
