// Developer tool (GPU box): what does a PLAIN read of the attention kernels' working set reach?
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/stream_ceiling tools/stream_ceiling.cu
//   /tmp/stream_ceiling > profiles/r02_stream_ceiling.json
//
// attention_fwd4 / attention_bwd4 read p_att16 + att_e16 = 2 x NL x 512 bf16 (116 MB at the bench's
// 56.8 K regions) once per launch, cold (the rest of the decode step -- ~100 MB of logits, gate
// pre-activations and weights -- has been through the 126 MB L2 in between).  This program times the
// cheapest possible kernel over the same bytes -- 16-byte loads, a register XOR, one store per thread
// -- launched like the attention kernels (one CTA of 512 threads per SM) and as a wide grid, with L2
// flushed by a 512 MB read before every launch, plus the same kernel over 2 GB (the asymptotic
// rate MEASURED_PEAKS.json's copy figure corresponds to).  The attention kernels' GB/s is to be
// read against the 116 MB lines, not against the 2 GB one: a ~25 us kernel pays launch, ramp-up and
// tail (the last CTA's last wave) out of its own time.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void read_kernel(const uint4* __restrict__ a, int64_t n16, uint32_t* __restrict__ out) {
  uint32_t acc = 0;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread and iteration
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 x0 = a[i], x1 = a[i + stride], x2 = a[i + 2 * stride], x3 = a[i + 3 * stride];
    acc ^= x0.x ^ x0.y ^ x0.z ^ x0.w ^ x1.x ^ x1.y ^ x1.z ^ x1.w ^ x2.x ^ x2.y ^ x2.z ^ x2.w ^ x3.x ^ x3.y ^ x3.z ^ x3.w;
  }
  for (; i < n16; i += stride) { const uint4 x = a[i]; acc ^= x.x ^ x.y ^ x.z ^ x.w; }
  if (acc == 0x12345678u) out[0] = acc;     // keeps the loads alive, (almost) never taken
}

__global__ void fill_kernel(uint4* a, int64_t n16, uint32_t v) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride)
    a[i] = make_uint4(v + uint32_t(i), v, v ^ uint32_t(i), v);
}

static double timed(const uint4* buf, int64_t bytes, int grid, int block, uint4* flush, int64_t flush_bytes,
                    uint32_t* out, int reps, bool do_flush) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int r = 0; r < reps + 2; ++r) {
    // flush with READS: a write flush leaves the L2 full of dirty lines whose write-back the timed
    // kernel then pays for (the first version of this tool did that and read 116 MB in 39 us)
    if (do_flush) read_kernel<<<1184, 512>>>(flush, flush_bytes / 16, out);
    CK(cudaEventRecord(e0));
    read_kernel<<<grid, block>>>(buf, bytes / 16, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t;
    CK(cudaEventElapsedTime(&t, e0, e1));
    if (r >= 2) ms.push_back(t);
  }
  std::sort(ms.begin(), ms.end());
  return ms[ms.size() / 2];
}

int main() {
  const int64_t NL = 56792;                      // regions of the bench batch (1024 rows, 10-100)
  const int64_t small = 2 * NL * 512 * 2;        // p_att16 + att_e16
  const int64_t big = int64_t(2) << 30;
  const int64_t flush_bytes = int64_t(512) << 20;
  uint4 *buf, *flush;
  uint32_t* out;
  CK(cudaMalloc(&buf, big));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMalloc(&out, 64));
  fill_kernel<<<1184, 512>>>(buf, big / 16, 1u);
  fill_kernel<<<1184, 512>>>(flush, flush_bytes / 16, 3u);
  CK(cudaDeviceSynchronize());
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  struct Case { const char* name; int64_t bytes; int grid, block; bool flush; };
  const Case cases[] = {
      {"116MB_cold_one_cta_per_sm_512thr", small, sms, 512, true},
      {"116MB_cold_two_ctas_per_sm_512thr", small, 2 * sms, 512, true},
      {"116MB_cold_four_ctas_per_sm_512thr", small, 4 * sms, 512, true},
      {"116MB_cold_wide_grid_256thr", small, 32 * sms, 256, true},
      {"116MB_l2_warm_four_ctas_per_sm", small, 4 * sms, 512, false},
      {"2GB_four_ctas_per_sm_512thr", big, 4 * sms, 512, true},
  };
  printf("{\n \"what\": \"plain 16-byte-load read kernel, median of 20 launches, CUDA events around the launch; L2 flushed by a 512 MB read of another buffer before every cold launch\",\n \"sms\": %d,\n \"cases\": {\n", sms);
  const int ncase = int(sizeof(cases) / sizeof(cases[0]));
  for (int i = 0; i < ncase; ++i) {
    const Case& c = cases[i];
    const double ms = timed(buf, c.bytes, c.grid, c.block, flush, flush_bytes, out, 20, c.flush);
    printf("  \"%s\": {\"bytes\": %lld, \"us\": %.2f, \"GBps\": %.1f}%s\n", c.name, (long long)c.bytes, ms * 1e3,
           double(c.bytes) / (ms * 1e-3) / 1e9, i + 1 < ncase ? "," : "");
  }
  printf(" }\n}\n");
  return 0;
}
