// CIDEr-D self-critical reward on the device (coopcap_cider_reward, include/coopcap.h).
//
// Reference: misc/rewards.py:34-71 turns the sampled / greedy id captions and every image's
// ground-truth captions into strings, and cider/pyciderevalcap/ciderD/ciderD_scorer.py scores them
// with Python dictionaries keyed by word tuples (float64) -- on the host, every training step,
// after two device->host copies.  Here an n-gram of ids is one exact 64-bit key, a caption's
// n-gram table is 64 slots, and the whole reward is five small launches that never leave the GPU.
// The arithmetic is integer/key matching plus a few hundred float64 operations per caption pair:
// latency-bound, nowhere near any roofline; what matters is that no host round trip remains.
//
// Every floating-point sum runs in the reference's order (dictionary insertion order = n-gram
// order k = 1..4, first occurrence by position; references in list order), so the scores differ
// from the reference only by the rounding of log / pow.
#include "../../include/coopcap.h"
#include "common.cuh"

namespace coopcap {

constexpr int CD_SLOTS = 64;      // 4 orders x 16 positions
constexpr int CD_PER = 16;
constexpr int CD_WARPS = 4;       // warps (captions) per block

__host__ __device__ inline uint64_t cider_mix(uint64_t x) {   // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__device__ __forceinline__ float df_lookup(const uint64_t* __restrict__ keys, const float* __restrict__ val,
                                           int cap, uint64_t key) {
  uint32_t h = uint32_t(cider_mix(key)) & uint32_t(cap - 1);
  for (int probe = 0; probe < cap; ++probe) {
    const uint64_t k = keys[h];
    if (k == key) return val[h];
    if (k == 0) return 0.f;                  // unseen n-gram: df 0 (defaultdict, ciderD_scorer.py:134)
    h = (h + 1) & uint32_t(cap - 1);
  }
  return 0.f;
}

__global__ void cider_img_rows_kernel(const int* __restrict__ row_img, int B, int n_img, int* __restrict__ img_rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    const int i = row_img[b];
    if (i >= 0 && i < n_img) atomicAdd(&img_rows[i], 1);
  }
}

// precook (ciderD_scorer.py:13-28) of every caption: one warp each
__global__ void __launch_bounds__(32 * CD_WARPS)
cider_cook_kernel(const int64_t* __restrict__ hyp0, const int64_t* __restrict__ hyp1,
                  const int64_t* __restrict__ refs, int B, int n_sets, int T, int W, int C_total,
                  uint64_t* __restrict__ ng_key, int* __restrict__ ng_cnt, int* __restrict__ ng_n,
                  int* __restrict__ ng_len) {
  const int c = blockIdx.x * CD_WARPS + (threadIdx.x >> 5);
  if (c >= C_total) return;
  const int l = threadIdx.x & 31;
  const int nh = n_sets * B;
  int width;
  int64_t w = 0;
  if (c < nh) {
    const int64_t* h = (c < B) ? hyp0 : hyp1;
    const int b = c % B;
    width = T;
    if (l < T) w = h[int64_t(l) * B + b];
  } else {
    width = W;
    if (l < W) w = refs[int64_t(c - nh) * W + l];
  }
  // words up to and including the first 0 (rewards.py:26-32)
  const unsigned zero = __ballot_sync(0xffffffffu, l < width && w <= 0);
  const int nw = zero ? (__ffs(zero)) : width;
  const uint64_t id = (uint64_t(w > 0 ? w : 0) + 1) & 0xffffull;
  uint64_t key = 0;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const uint64_t nxt = __shfl_down_sync(0xffffffffu, id, n);
    key |= nxt << (16 * n);
    const int npos = nw - n;
    const bool valid = l < npos;
    int cnt = 0;
    bool first = valid;
    for (int j = 0; j < npos; ++j) {
      const uint64_t kj = __shfl_sync(0xffffffffu, key, j);
      if (valid && kj == key) {
        ++cnt;
        if (j < l) first = false;
      }
    }
    const unsigned firsts = __ballot_sync(0xffffffffu, first);
    if (first) {
      const int slot = n * CD_PER + __popc(firsts & ((1u << l) - 1u));
      ng_key[int64_t(c) * CD_SLOTS + slot] = key;
      ng_cnt[int64_t(c) * CD_SLOTS + slot] = cnt;
    }
    if (l == 0) ng_n[c * 4 + n] = __popc(firsts);
  }
  if (l == 0) ng_len[c] = nw > 1 ? nw - 1 : 0;     // sum of bigram counts (ciderD_scorer.py:143-144)
}

// compute_doc_freq (ciderD_scorer.py:105-118): every entry's reference set counts an n-gram once;
// an image serves img_rows * n_sets entries
__global__ void __launch_bounds__(256)
cider_df_kernel(const uint64_t* __restrict__ ng_key, const int* __restrict__ ng_n,
                const int* __restrict__ ref_off, const int* __restrict__ img_rows, int n_img, int n_ref,
                int nh, int n_sets, uint64_t* __restrict__ df_keys, float* __restrict__ df_val, int cap) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_ref * CD_SLOTS) return;
  const int r = idx / CD_SLOTS, slot = idx % CD_SLOTS, n = slot / CD_PER, j = slot % CD_PER;
  const int c = nh + r;
  if (j >= ng_n[c * 4 + n]) return;
  int lo = 0, hi = n_img;                      // image of reference r: last i with ref_off[i] <= r
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ref_off[mid] <= r) lo = mid; else hi = mid;
  }
  const int img = lo;
  const int weight = img_rows[img] * n_sets;
  if (weight == 0) return;
  const uint64_t key = ng_key[int64_t(c) * CD_SLOTS + slot];
  for (int r2 = ref_off[img]; r2 < r; ++r2) {  // already counted by an earlier caption of the image?
    const int c2 = nh + r2, m = ng_n[c2 * 4 + n];
    for (int q = 0; q < m; ++q)
      if (ng_key[int64_t(c2) * CD_SLOTS + n * CD_PER + q] == key) return;
  }
  uint32_t h = uint32_t(cider_mix(key)) & uint32_t(cap - 1);
  for (int probe = 0; probe < cap; ++probe) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(df_keys + h), 0ull,
                                              static_cast<unsigned long long>(key));
    if (prev == 0ull || prev == key) {
      atomicAdd(df_val + h, float(weight));
      return;
    }
    h = (h + 1) & uint32_t(cap - 1);
  }
}

// counts2vec (ciderD_scorer.py:121-145)
__global__ void __launch_bounds__(32 * CD_WARPS)
cider_vec_kernel(const uint64_t* __restrict__ ng_key, const int* __restrict__ ng_cnt,
                 const int* __restrict__ ng_n, int C_total, const uint64_t* __restrict__ df_keys,
                 const float* __restrict__ df_val, int cap, double log_ref_len,
                 double* __restrict__ ng_w, double* __restrict__ ng_norm) {
  __shared__ double s_w[CD_WARPS][CD_SLOTS];
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int c = blockIdx.x * CD_WARPS + wi;
  if (c >= C_total) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int slot = l + 32 * half, n = slot / CD_PER, j = slot % CD_PER;
    double w = 0.0;
    if (j < ng_n[c * 4 + n]) {
      const float df = df_lookup(df_keys, df_val, cap, ng_key[int64_t(c) * CD_SLOTS + slot]);
      w = double(ng_cnt[int64_t(c) * CD_SLOTS + slot]) * (log_ref_len - log(fmax(1.0, double(df))));
    }
    s_w[wi][slot] = w;
    ng_w[int64_t(c) * CD_SLOTS + slot] = w;
  }
  __syncwarp();
  if (l < 4) {
    double acc = 0.0;
    const int m = ng_n[c * 4 + l];
    for (int j = 0; j < m; ++j) acc += s_w[wi][l * CD_PER + j] * s_w[wi][l * CD_PER + j];
    ng_norm[c * 4 + l] = sqrt(acc);
  }
}

// sim + the per-hypothesis average (ciderD_scorer.py:147-203): one warp per hypothesis
__global__ void __launch_bounds__(32 * CD_WARPS)
cider_score_kernel(const uint64_t* __restrict__ ng_key, const int* __restrict__ ng_n,
                   const int* __restrict__ ng_len, const double* __restrict__ ng_w,
                   const double* __restrict__ ng_norm, const int* __restrict__ ref_off,
                   const int* __restrict__ row_img, int B, int nh, int n_img,
                   double* __restrict__ scores) {
  __shared__ uint64_t s_key[CD_WARPS][CD_SLOTS];
  __shared__ double s_wr[CD_WARPS][CD_SLOTS];
  __shared__ double s_term[CD_WARPS][CD_SLOTS];
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int c = blockIdx.x * CD_WARPS + wi;
  if (c >= nh) return;
  const int img = row_img[c % B];
  const bool ok = img >= 0 && img < n_img;
  const int r0 = ok ? ref_off[img] : 0, r1 = ok ? ref_off[img + 1] : 0;
  uint64_t hk[2];
  double hw[2];
  bool hv[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int slot = l + 32 * half, n = slot / CD_PER, j = slot % CD_PER;
    hv[half] = j < ng_n[c * 4 + n];
    hk[half] = ng_key[int64_t(c) * CD_SLOTS + slot];
    hw[half] = ng_w[int64_t(c) * CD_SLOTS + slot];
  }
  const int len_h = ng_len[c];
  const double norm_h = l < 4 ? ng_norm[c * 4 + l] : 0.0;
  const int m_h = l < 4 ? ng_n[c * 4 + l] : 0;
  double score = 0.0;                               // lane n < 4: running score of order n
  for (int r = r0; r < r1; ++r) {
    const int cr = nh + r;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int slot = l + 32 * half;
      s_key[wi][slot] = ng_key[int64_t(cr) * CD_SLOTS + slot];
      s_wr[wi][slot] = ng_w[int64_t(cr) * CD_SLOTS + slot];
    }
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int slot = l + 32 * half, n = slot / CD_PER;
      double term = 0.0;
      if (hv[half]) {
        const int m = ng_n[cr * 4 + n];
        for (int q = 0; q < m; ++q)
          if (s_key[wi][n * CD_PER + q] == hk[half]) {
            const double wr = s_wr[wi][n * CD_PER + q];
            term = fmin(hw[half], wr) * wr;          // clipped product (ciderD_scorer.py:166)
            break;
          }
      }
      s_term[wi][slot] = term;
    }
    __syncwarp();
    if (l < 4) {
      double val = 0.0;
      for (int j = 0; j < m_h; ++j) val += s_term[wi][l * CD_PER + j];
      const double norm_r = ng_norm[cr * 4 + l];
      if (norm_h != 0.0 && norm_r != 0.0) val /= (norm_h * norm_r);
      const double delta = double(len_h - ng_len[cr]);
      val *= pow(2.718281828459045, -(delta * delta) / (2.0 * 6.0 * 6.0));   // sigma = 6
      score += val;
    }
    __syncwarp();
  }
  const double s1 = __shfl_sync(0xffffffffu, score, 1), s2 = __shfl_sync(0xffffffffu, score, 2),
               s3 = __shfl_sync(0xffffffffu, score, 3);
  if (l == 0) {
    double s = (((score + s1) + s2) + s3) / 4.0;    // np.mean over the four orders
    const int nref = r1 - r0;
    s = nref > 0 ? s / double(nref) : 0.0;
    scores[c] = s * 10.0;
  }
}

// reward, REINFORCE coefficients and the logged statistics
__global__ void __launch_bounds__(1024)
cider_finish_kernel(const double* __restrict__ scores, const int64_t* __restrict__ hyp0, int B, int T,
                    int n_sets, int differenced, float* __restrict__ reward, float* __restrict__ coef,
                    double* __restrict__ stats) {
  __shared__ int s_n;
  __shared__ double s_acc[4];
  if (threadIdx.x == 0) { s_n = 0; s_acc[0] = s_acc[1] = s_acc[2] = s_acc[3] = 0.0; }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int k = T;
    for (int t = T - 1; t >= 0; --t)
      if (hyp0[int64_t(t) * B + b] <= 0) k = t;
    atomicMax(&s_n, k);
  }
  __syncthreads();
  const int n = s_n;                                 // caption width of the sampled set
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int k = T;
    for (int t = T - 1; t >= 0; --t)
      if (hyp0[int64_t(t) * B + b] <= 0) k = t;
    const double gen = scores[b], gr = n_sets > 1 ? scores[B + b] : 0.0;
    const double r = (differenced && n_sets > 1) ? gen - gr : gen;
    reward[b] = float(r);
    atomicAdd(&s_acc[0], r);
    atomicAdd(&s_acc[1], gr);
    atomicAdd(&s_acc[2], double(min(k + 1, n)));
    atomicAdd(&s_acc[3], gen);
  }
  __syncthreads();
  const double msum = s_acc[2];
  if (coef) {
    const float inv = msum > 0.0 ? float(1.0 / msum) : 0.f;
    for (int i = threadIdx.x; i < T * B; i += blockDim.x) {
      const int t = i / B, b = i % B;
      int k = T;
      for (int u = T - 1; u >= 0; --u)
        if (hyp0[int64_t(u) * B + b] <= 0) k = u;
      coef[i] = (t < min(k + 1, n)) ? -reward[b] * inv : 0.f;
    }
  }
  if (threadIdx.x == 0) {
    stats[0] = s_acc[0] / B;
    stats[1] = s_acc[1] / B;
    stats[2] = msum;
    stats[3] = s_acc[3] / B;
  }
}

}  // namespace coopcap

extern "C" {

uint64_t coopcap_cider_hash(uint64_t key) { return coopcap::cider_mix(key); }

int coopcap_cider_reward(const coopcap_cider* c, coopcap_stream_t stream) {
  using namespace coopcap;
  CC_REQUIRE(c != nullptr, "cider: null context");
  CC_REQUIRE(c->B > 0 && (c->n_sets == 1 || c->n_sets == 2) && c->n_img > 0 && c->n_ref > 0,
             "cider: bad sizes B=%d n_sets=%d n_img=%d n_ref=%d", c->B, c->n_sets, c->n_img, c->n_ref);
  CC_REQUIRE(c->T > 0 && c->T <= CD_PER && c->W > 0 && c->W <= CD_PER,
             "cider: captions of more than %d ids are not supported (T=%d W=%d)", CD_PER, c->T, c->W);
  CC_REQUIRE(c->hyp0 && (c->n_sets == 1 || c->hyp1) && c->refs && c->ref_off && c->row_img,
             "cider: null caption input");
  CC_REQUIRE(c->df_keys && c->df_val && c->df_cap > 0 && (c->df_cap & (c->df_cap - 1)) == 0,
             "cider: df table missing or its capacity %d is not a power of two", c->df_cap);
  CC_REQUIRE(!c->corpus || int64_t(c->df_cap) >= 2 * 58 * int64_t(c->n_ref),
             "cider: corpus-mode table of %d slots is too small for %d references", c->df_cap, c->n_ref);
  CC_REQUIRE(c->ng_key && c->ng_cnt && c->ng_n && c->ng_len && c->ng_w && c->ng_norm && c->img_rows,
             "cider: workspace missing");
  CC_REQUIRE(c->scores && c->reward && c->stats, "cider: output missing");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int nh = c->n_sets * c->B, C_total = nh + c->n_ref;
  const int cap_blocks = (C_total + CD_WARPS - 1) / CD_WARPS;
  cider_cook_kernel<<<cap_blocks, 32 * CD_WARPS, 0, s>>>(c->hyp0, c->hyp1, c->refs, c->B, c->n_sets, c->T,
                                                        c->W, C_total, c->ng_key, c->ng_cnt, c->ng_n,
                                                        c->ng_len);
  CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  double log_ref_len = c->log_ref_len;
  if (c->corpus) {
    log_ref_len = log(double(nh));                   // ciderD_scorer.py:178-179
    CC_CHECK_CUDA(cudaMemsetAsync(c->img_rows, 0, sizeof(int) * c->n_img, s));
    CC_CHECK_CUDA(cudaMemsetAsync(c->df_keys, 0, sizeof(uint64_t) * c->df_cap, s));
    CC_CHECK_CUDA(cudaMemsetAsync(c->df_val, 0, sizeof(float) * c->df_cap, s));
    cider_img_rows_kernel<<<(c->B + 255) / 256, 256, 0, s>>>(c->row_img, c->B, c->n_img, c->img_rows);
    CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
    const int n = c->n_ref * CD_SLOTS;
    cider_df_kernel<<<(n + 255) / 256, 256, 0, s>>>(c->ng_key, c->ng_n, c->ref_off, c->img_rows, c->n_img,
                                                   c->n_ref, nh, c->n_sets, c->df_keys, c->df_val, c->df_cap);
    CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  }
  cider_vec_kernel<<<cap_blocks, 32 * CD_WARPS, 0, s>>>(c->ng_key, c->ng_cnt, c->ng_n, C_total, c->df_keys,
                                                       c->df_val, c->df_cap, log_ref_len, c->ng_w,
                                                       c->ng_norm);
  CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  cider_score_kernel<<<(nh + CD_WARPS - 1) / CD_WARPS, 32 * CD_WARPS, 0, s>>>(
      c->ng_key, c->ng_n, c->ng_len, c->ng_w, c->ng_norm, c->ref_off, c->row_img, c->B, nh, c->n_img,
      c->scores);
  CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  cider_finish_kernel<<<1, 1024, 0, s>>>(c->scores, c->hyp0, c->B, c->T, c->n_sets, c->differenced, c->reward,
                                         c->coef, c->stats);
  CC_LAUNCH_CHECK_K(PROF_MISC, s, 0.0, 0.0);
  return CC_OK;
}

}  // extern "C"
