// Host-side packing of the region features: fp32 [B, L, D] (padded, host memory) -> bf16 [NL, D]
// (valid regions only, pinned staging buffer), done by a small persistent pool of worker threads so
// that only a quarter of the reference's bytes (839 MB padded fp32 -> ~230 MB packed bf16 at 1024
// rows) cross PCIe, by DMA, without occupying any SM.  Same rounding (round-to-nearest-even) as the
// device-side pack kernel, so both upload paths produce bit-identical operands.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif
#include "common.cuh"
#include "coopcap.h"

namespace coopcap {
namespace {

inline uint16_t bf16_rne(uint32_t u) {
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40u);   // quiet NaN
  return uint16_t((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

void convert_scalar(const float* src, uint16_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    uint32_t u;
    std::memcpy(&u, src + i, 4);
    dst[i] = bf16_rne(u);
  }
}

#if defined(__x86_64__)
// 16 floats per iteration; dst 32-byte aligned -> non-temporal stores (the staging buffer is only
// read by the DMA engine afterwards).  NaNs take the scalar path of their 16-element group.
__attribute__((target("avx2"))) void convert_avx2(const float* src, uint16_t* dst, size_t n) {
  const __m256i one = _mm256_set1_epi32(1), bias = _mm256_set1_epi32(0x7fff);
  const __m256i absmask = _mm256_set1_epi32(0x7fffffff), inf = _mm256_set1_epi32(0x7f800000);
  size_t i = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
    const __m256i nan = _mm256_or_si256(_mm256_cmpgt_epi32(_mm256_and_si256(a, absmask), inf),
                                        _mm256_cmpgt_epi32(_mm256_and_si256(b, absmask), inf));
    if (!_mm256_testz_si256(nan, nan)) {
      convert_scalar(src + i, dst + i, 16);
      continue;
    }
    const __m256i ra = _mm256_srli_epi32(
        _mm256_add_epi32(a, _mm256_add_epi32(bias, _mm256_and_si256(_mm256_srli_epi32(a, 16), one))), 16);
    const __m256i rb = _mm256_srli_epi32(
        _mm256_add_epi32(b, _mm256_add_epi32(bias, _mm256_and_si256(_mm256_srli_epi32(b, 16), one))), 16);
    const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(ra, rb), 0xD8);
    if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), p);
    else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), p);
  }
  convert_scalar(src + i, dst + i, n - i);
  _mm_sfence();
}
#endif

void convert(const float* src, uint16_t* dst, size_t n) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return convert_avx2(src, dst, n);
#endif
  convert_scalar(src, dst, n);
}

struct Job {
  int id;
  const float* src;
  const int* off;        // [B+1] or null
  int B, L, D;
  uint16_t* dst;
  std::atomic<int> next{0};      // next row to claim
  int done = 0;                  // rows finished        (guarded by the pool mutex)
  int users = 0;                 // workers holding a pointer to this job (guarded by the pool mutex)
};

class Pool {
 public:
  static Pool& get() {
    static Pool p;
    return p;
  }
  int start(const float* src, const int* off, int B, int L, int D, void* dst, int nthreads) {
    std::unique_lock<std::mutex> lk(mu_);
    grow(nthreads);
    auto* j = new Job();
    j->id = next_id_++;
    j->src = src; j->off = off; j->B = B; j->L = L; j->D = D;
    j->dst = reinterpret_cast<uint16_t*>(dst);
    jobs_.push_back(j);
    cv_work_.notify_all();
    return j->id;
  }
  int wait(int id) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] {
      for (Job* j : jobs_)
        if (j->id == id) return false;
      return true;
    });
    return 0;
  }

 private:
  Pool() = default;
  ~Pool() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
      cv_work_.notify_all();
    }
    for (auto& t : threads_) t.join();
  }
  void grow(int n) {
    if (n <= 0) {
      n = int(std::thread::hardware_concurrency()) - 1;   // leave the enqueueing thread its core
      if (n < 1) n = 1;
    }
    if (n > 64) n = 64;
    while (int(threads_.size()) < n) {
      threads_.emplace_back([this] { run(); });
#if defined(__linux__)
      // workers stay off core 0 so that the thread enqueueing kernels keeps a core of its own
      const unsigned ncpu = std::thread::hardware_concurrency();
      if (ncpu > 1) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(1 + (threads_.size() - 1) % (ncpu - 1), &set);
        pthread_setaffinity_np(threads_.back().native_handle(), sizeof(set), &set);
      }
#endif
    }
  }
  void run() {
    for (;;) {
      Job* j = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] {
          if (stop_) return true;
          for (Job* q : jobs_)
            if (q->next.load(std::memory_order_relaxed) < q->B) return true;
          return false;
        });
        if (stop_) return;
        for (Job* q : jobs_)
          if (q->next.load(std::memory_order_relaxed) < q->B) { j = q; break; }
        if (j) ++j->users;
      }
      if (!j) continue;
      int finished = 0;
      for (;;) {
        const int b = j->next.fetch_add(1);
        if (b >= j->B) break;
        const int r0 = j->off ? j->off[b] : b * j->L;
        const int len = j->off ? j->off[b + 1] - r0 : j->L;
        convert(j->src + size_t(b) * j->L * j->D, j->dst + size_t(r0) * j->D, size_t(len) * j->D);
        ++finished;
      }
      {
        std::unique_lock<std::mutex> lk(mu_);
        j->done += finished;
        --j->users;
        if (j->done == j->B && j->users == 0) {     // the last worker out retires the job
          for (auto it = jobs_.begin(); it != jobs_.end(); ++it)
            if (*it == j) { jobs_.erase(it); break; }
          delete j;
          cv_done_.notify_all();
        }
      }
    }
  }
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  std::deque<Job*> jobs_;
  std::vector<std::thread> threads_;
  int next_id_ = 0;
  bool stop_ = false;
};

}  // namespace
}  // namespace coopcap

extern "C" {

int coopcap_host_pack_start(const float* att_feats, const int* att_off_host, int B, int L, int D,
                            void* att16_host, int nthreads) {
  using namespace coopcap;
  CC_REQUIRE(att_feats && att16_host && B > 0 && L > 0 && D > 0, "host_pack: bad arguments");
  if (att_off_host) {
    CC_REQUIRE(att_off_host[0] == 0, "host_pack: att_off[0] must be 0");
    for (int b = 0; b < B; ++b)
      CC_REQUIRE(att_off_host[b + 1] >= att_off_host[b] && att_off_host[b + 1] - att_off_host[b] <= L,
                 "host_pack: row %d has %d regions (L = %d)", b, att_off_host[b + 1] - att_off_host[b], L);
  }
  return Pool::get().start(att_feats, att_off_host, B, L, D, att16_host, nthreads);
}

int coopcap_host_pack_wait(int job) {
  using namespace coopcap;
  CC_REQUIRE(job >= 0, "host_pack_wait: bad job id %d", job);
  return Pool::get().wait(job);
}

}  // extern "C"
