// Batched beam search (coopcap_speaker_beam_fwd) and retrieval ranks (coopcap_retrieval_ranks).
//
// Reference: AttModel.sample_beam (models/AttModel.py:150-289) decodes ONE image at a time:
// beam_size rows through the core, the [beam, V+1] log-probabilities copied to the host, a CPU
// sort, a Python list of beam^2 candidates, per-slot state copies; eval_utils.i2t / t2i
// (eval_utils.py:545-720) argsort one score row per query in numpy.  Here every image advances in
// the same launches: rows are beam-major (row = slot * n_img + image), so the per-step kernels of
// the decode loop (gate GEMM, attention, a2c GEMM, LSTM pointwise, logit GEMM) are reused
// unchanged -- the attention kernel runs once per slot over the n_img images -- and one CTA per
// image merges the candidates exactly as the reference does (oracle/speaker.py `sample_beam` lists
// the quirks that are kept).
#include <algorithm>
#include "../../include/coopcap.h"
#include "common.cuh"
#include "gemm.cuh"
#include "speaker_kernels.cuh"

namespace coopcap {

using bf16 = __nv_bfloat16;

int attention_fwd_launch(const coopcap_speaker* c, const float* s_t, bf16* att_res16_t, float* att_w_t,
                         cudaStream_t s, float* att_res32_t);
int lstm_fwd_launch(const coopcap_speaker* c, const float* s_t, const float* u_t, const float* c_prev,
                    float* c_next, bf16* h_out16, int64_t ld_h, bf16* out16_t, const uint8_t* keep,
                    uint64_t site, float drop_p, int rows, cudaStream_t s);

constexpr int BEAM_MAX = 8;
constexpr int BEAM_THREADS = 256;

// x_0 = relu(embed[BOS]), h = c = 0 for every row
__global__ void beam_start_kernel(const float* __restrict__ embed, int64_t bos, int E, int R,
                                  bf16* __restrict__ xh0, float* __restrict__ c0) {
  const int r = blockIdx.x;
  bf16* row = xh0 + int64_t(r) * (E + R);
  embed_row(embed, bos, E, nullptr, 0, 0, 0, 0.f, row);
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    row[E + i] = __float2bfloat16_rn(0.f);
    c0[int64_t(r) * R + i] = 0.f;
  }
}

// (value, index) maximum of the block, ties -> the smaller index; every thread gets the result
__device__ __forceinline__ void block_argmax(float& v, int& i, float* s_v, int* s_i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = v; s_i[threadIdx.x >> 5] = i; }
  __syncthreads();
  v = s_v[0]; i = s_i[0];
#pragma unroll
  for (int w = 1; w < BEAM_THREADS / 32; ++w)
    if (s_v[w] > v || (s_v[w] == v && s_i[w] < i)) { v = s_v[w]; i = s_i[w]; }
}

// One merge step (AttModel.py:196-257) for one image per CTA.  step = t in the reference (1..T).
__global__ void __launch_bounds__(BEAM_THREADS)
beam_step_kernel(const float* __restrict__ logits, int V1, int n_img, int bs, int step, int T, int no_repeat,
                 const int64_t* __restrict__ hseq_src, const float* __restrict__ hlp_src,
                 int64_t* __restrict__ hseq_dst, float* __restrict__ hlp_dst, float* __restrict__ beam_sum,
                 const int* __restrict__ forced_parent, const int64_t* __restrict__ forced_tok,
                 int* __restrict__ raw_parent, int64_t* __restrict__ raw_tok, int* __restrict__ parent,
                 int64_t* __restrict__ tok, int64_t* __restrict__ done_seq, float* __restrict__ done_lp,
                 int* __restrict__ done_slot, float* __restrict__ done_p_rec, int* __restrict__ done_n) {
  __shared__ float s_v[BEAM_THREADS / 32];
  __shared__ int s_i[BEAM_THREADS / 32];
  __shared__ float s_red[8];
  __shared__ float s_ys[BEAM_MAX][BEAM_MAX];   // [slot][word rank] log-probability
  __shared__ int s_ix[BEAM_MAX][BEAM_MAX];
  __shared__ float s_lse[BEAM_MAX];
  __shared__ int s_par[BEAM_MAX];
  __shared__ int s_tok[BEAM_MAX];
  __shared__ float s_r[BEAM_MAX], s_p[BEAM_MAX];
  const int b = blockIdx.x;
  const int rows = step == 1 ? 1 : bs;            // first merge: only slot 0 is live (:208-210)
  const int cols = min(bs, V1);
  const int64_t hbase = int64_t(b) * T * bs;
  for (int q = 0; q < rows; ++q) {
    const float* z = logits + (int64_t(q) * n_img + b) * V1;
    const int ban = (no_repeat && step > 1) ? int(hseq_src[hbase + int64_t(step - 2) * bs + q]) : -1;
    // log-sum-exp over the whole row (the constraint is added AFTER log_softmax, :204-207)
    float m = -INFINITY;
    for (int v = threadIdx.x; v < V1; v += BEAM_THREADS) m = fmaxf(m, z[v]);
    m = block_max_256(m, s_red);
    float sum = 0.f;
    for (int v = threadIdx.x; v < V1; v += BEAM_THREADS) sum += __expf(z[v] - m);
    sum = block_sum_256(sum, s_red);
    const float lse = m + logf(sum);
    // this thread's own top `cols` (descending, ties -> smaller index), then `cols` block rounds
    float tv[BEAM_MAX];
    int ti[BEAM_MAX];
#pragma unroll
    for (int k = 0; k < BEAM_MAX; ++k) { tv[k] = -INFINITY; ti[k] = 0x7fffffff; }
    for (int v = threadIdx.x; v < V1; v += BEAM_THREADS) {
      float x = (v == ban) ? -INFINITY : z[v];
      int xi = v;
#pragma unroll
      for (int k = 0; k < BEAM_MAX; ++k) {
        if (k < cols && (x > tv[k] || (x == tv[k] && xi < ti[k]))) {
          const float fv = tv[k]; const int fi = ti[k];
          tv[k] = x; ti[k] = xi; x = fv; xi = fi;
        }
      }
    }
    for (int cc = 0; cc < cols; ++cc) {
      float v = tv[0];
      int i = ti[0];
      block_argmax(v, i, s_v, s_i);
      if (ti[0] == i) {                            // the winner pops its head
#pragma unroll
        for (int k = 0; k + 1 < BEAM_MAX; ++k) { tv[k] = tv[k + 1]; ti[k] = ti[k + 1]; }
        tv[BEAM_MAX - 1] = -INFINITY; ti[BEAM_MAX - 1] = 0x7fffffff;
      }
      if (threadIdx.x == 0) { s_ys[q][cc] = v - lse; s_ix[q][cc] = i; }
    }
    if (threadIdx.x == 0) s_lse[q] = lse;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // candidates word-rank-major, stable insertion sort by total log-probability (:211-221)
    float cp[BEAM_MAX * BEAM_MAX], cr[BEAM_MAX * BEAM_MAX];
    int cq[BEAM_MAX * BEAM_MAX], cw[BEAM_MAX * BEAM_MAX];
    int n = 0;
    for (int cc = 0; cc < cols; ++cc)
      for (int q = 0; q < rows; ++q) {
        const float r = s_ys[q][cc], p = beam_sum[b * bs + q] + r;
        int pos = n;
        while (pos > 0 && p > cp[pos - 1]) {       // strict: equal scores keep their order
          cp[pos] = cp[pos - 1]; cr[pos] = cr[pos - 1]; cq[pos] = cq[pos - 1]; cw[pos] = cw[pos - 1];
          --pos;
        }
        cp[pos] = p; cr[pos] = r; cq[pos] = q; cw[pos] = s_ix[q][cc];
        ++n;
      }
    for (int vix = 0; vix < bs; ++vix) {
      // fewer candidates than slots only when beam_size > V1 (the reference asserts it away)
      const int k = vix < n ? vix : n - 1;
      int q = cq[k], w = cw[k];
      float r = cr[k], p = cp[k];
      if (raw_parent) raw_parent[(int64_t(step - 1) * n_img + b) * bs + vix] = q;
      if (raw_tok) raw_tok[(int64_t(step - 1) * n_img + b) * bs + vix] = w;
      if (forced_parent && forced_tok) {
        q = forced_parent[(int64_t(step - 1) * n_img + b) * bs + vix];
        w = int(forced_tok[(int64_t(step - 1) * n_img + b) * bs + vix]);
        const int ban = (no_repeat && step > 1) ? int(hseq_src[hbase + int64_t(step - 2) * bs + q]) : -1;
        r = (w == ban) ? -INFINITY : logits[(int64_t(q) * n_img + b) * V1 + w] - s_lse[q];
        p = beam_sum[b * bs + q] + r;
      }
      s_par[vix] = q; s_tok[vix] = w; s_r[vix] = r; s_p[vix] = p;
    }
  }
  __syncthreads();
  // fork the histories (:224-246): dst column vix = src column parent, then the new word
  for (int i = threadIdx.x; i < (step - 1) * bs; i += BEAM_THREADS) {
    const int t = i / bs, vix = i % bs;
    hseq_dst[hbase + int64_t(t) * bs + vix] = hseq_src[hbase + int64_t(t) * bs + s_par[vix]];
    hlp_dst[hbase + int64_t(t) * bs + vix] = hlp_src[hbase + int64_t(t) * bs + s_par[vix]];
  }
  if (threadIdx.x < bs) {
    const int vix = threadIdx.x;
    hseq_dst[hbase + int64_t(step - 1) * bs + vix] = s_tok[vix];
    hlp_dst[hbase + int64_t(step - 1) * bs + vix] = s_r[vix];
    beam_sum[b * bs + vix] = s_p[vix];
    parent[vix * n_img + b] = s_par[vix];
    tok[vix * n_img + b] = s_tok[vix];
  }
  __syncthreads();
  // record finished beams (:247-255), slot order
  if (threadIdx.x == 0) {
    int nd = done_n[b];
    for (int vix = 0; vix < bs; ++vix)
      if (s_tok[vix] == 0 || step == T) {
        const int64_t e = int64_t(b) * bs * T + nd;
        for (int t = 0; t < T; ++t) {
          done_seq[e * T + t] = t < step ? hseq_dst[hbase + int64_t(t) * bs + vix] : 0;
          done_lp[e * T + t] = t < step ? hlp_dst[hbase + int64_t(t) * bs + vix] : 0.f;
        }
        done_slot[e] = vix;
        done_p_rec[e] = s_p[vix];
        ++nd;
      }
    done_n[b] = nd;
  }
}

// state of the new slots (:237-243) and their next input (:258-263, eval mode: no dropout)
__global__ void beam_reorder_kernel(const int* __restrict__ parent, const int64_t* __restrict__ tok,
                                    const bf16* __restrict__ h_stage, const float* __restrict__ c_stage,
                                    const float* __restrict__ embed, int E, int R, int n_img,
                                    bf16* __restrict__ xh_next, float* __restrict__ c_next) {
  const int r = blockIdx.x, b = r % n_img;
  const int src = parent[r] * n_img + b;
  bf16* row = xh_next + int64_t(r) * (E + R);
  embed_row(embed, tok[r], E, nullptr, 0, 0, 0, 0.f, row);
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    row[E + i] = h_stage[int64_t(src) * R + i];
    c_next[int64_t(r) * R + i] = c_stage[int64_t(src) * R + i];
  }
}

// rank the recorded beams (:283-287): score = the FINAL running sum of the slot an entry was recorded
// in (the reference stores a view of the running-sum tensor), ties -> recording order
__global__ void beam_finish_kernel(const int64_t* __restrict__ done_seq, const float* __restrict__ done_lp,
                                   const int* __restrict__ done_slot, const int* __restrict__ done_n,
                                   const float* __restrict__ beam_sum, int bs, int T,
                                   float* __restrict__ done_p, int64_t* __restrict__ seq,
                                   float* __restrict__ seq_logp) {
  const int b = blockIdx.x;
  __shared__ int s_best;
  if (threadIdx.x == 0) {
    int best = 0;
    float bp = -INFINITY;
    const int n = done_n[b];
    for (int e = 0; e < n; ++e) {
      const float p = beam_sum[b * bs + done_slot[int64_t(b) * bs * T + e]];
      done_p[int64_t(b) * bs * T + e] = p;
      if (e == 0 || p > bp) { bp = p; best = e; }
    }
    s_best = best;
  }
  __syncthreads();
  const int64_t e = int64_t(b) * bs * T + s_best;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    seq[int64_t(b) * T + t] = done_seq[e * T + t];
    seq_logp[int64_t(b) * T + t] = done_lp[e * T + t];
  }
}

int speaker_beam_fwd(const coopcap_speaker* c, const coopcap_beam* bm, cudaStream_t s) {
  CC_REQUIRE(c != nullptr && bm != nullptr, "beam: null context");
  CC_REQUIRE(bm->beam_size >= 1 && bm->beam_size <= BEAM_MAX, "beam: beam_size %d outside 1..%d",
             bm->beam_size, BEAM_MAX);
  CC_REQUIRE(bm->T >= 1 && c->B > 0 && c->NL > 0, "beam: bad sizes T=%d B=%d NL=%d", bm->T, c->B, c->NL);
  CC_REQUIRE(c->R % 8 == 0 && c->E % 8 == 0 && c->A % 8 == 0 && c->V1 >= bm->beam_size,
             "beam: R, E, A must be multiples of 8 and beam_size <= V+1 (AttModel.py:164)");
  CC_REQUIRE(bm->xh16 && bm->c2 && bm->s_t && bm->u_t && bm->att_res16 && bm->att_w && bm->h_stage16 &&
                 bm->c_stage && bm->logits && bm->parent && bm->tok && bm->hist_seq && bm->hist_lp &&
                 bm->beam_sum, "beam: workspace missing");
  CC_REQUIRE(bm->done_seq && bm->done_lp && bm->done_slot && bm->done_p_rec && bm->done_p && bm->done_n &&
                 bm->seq && bm->seq_logp, "beam: output missing");
  CC_REQUIRE((bm->forced_parent == nullptr) == (bm->forced_tok == nullptr),
             "beam: forced_parent and forced_tok come together");
  const int n_img = c->B, bs = bm->beam_size, T = bm->T, rows = n_img * bs;
  const int R = c->R, E = c->E, A = c->A, V1 = c->V1, NS = 5 * R + A, XH = E + R;
  bf16* xh16 = reinterpret_cast<bf16*>(bm->xh16);
  bf16* att_res16 = reinterpret_cast<bf16*>(bm->att_res16);
  bf16* h_stage16 = reinterpret_cast<bf16*>(bm->h_stage16);
  int rc;
  CC_CHECK_CUDA(cudaMemsetAsync(bm->hist_seq, 0, sizeof(int64_t) * 2 * n_img * T * bs, s));
  CC_CHECK_CUDA(cudaMemsetAsync(bm->hist_lp, 0, sizeof(float) * 2 * n_img * T * bs, s));
  CC_CHECK_CUDA(cudaMemsetAsync(bm->beam_sum, 0, sizeof(float) * n_img * bs, s));
  CC_CHECK_CUDA(cudaMemsetAsync(bm->done_n, 0, sizeof(int) * n_img, s));
  beam_start_kernel<<<rows, 128, 0, s>>>(c->embed, c->start_token, E, R, xh16, bm->c2);
  CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 0.0);
  for (int t = 0; t < T; ++t) {                    // core evaluation t feeds merge step t + 1
    bf16* xh_cur = xh16 + int64_t(t & 1) * rows * XH;
    bf16* xh_nxt = xh16 + int64_t((t + 1) & 1) * rows * XH;
    float* c_cur = bm->c2 + int64_t(t & 1) * rows * R;
    float* c_nxt = bm->c2 + int64_t((t + 1) & 1) * rows * R;
    EpiStoreParams e1 = {};
    e1.alpha = 1.f; e1.bias = c->b_cat; e1.C = bm->s_t; e1.ldc = NS;
    if ((rc = gemm_run(0, 0, 0, xh_cur, XH, c->w_cat16, XH, rows, NS, XH, 1, 0, e1, s))) return rc;
    for (int q = 0; q < bs; ++q)                   // slot q: the n_img images, the context's regions
      if ((rc = attention_fwd_launch(c, bm->s_t + int64_t(q) * n_img * NS, att_res16 + int64_t(q) * n_img * R,
                                     bm->att_w + int64_t(q) * c->NL, s, nullptr)))
        return rc;
    EpiStoreParams e2 = {};
    e2.alpha = 1.f; e2.bias = c->b_a2c; e2.C = bm->u_t; e2.ldc = 2 * R;
    if ((rc = gemm_run(0, 0, 0, att_res16, R, c->w_a2c16, R, rows, 2 * R, R, 1, 0, e2, s))) return rc;
    // h -> staging (re-ordered below); out16 = h in evaluation mode: written over att_res16
    if ((rc = lstm_fwd_launch(c, bm->s_t, bm->u_t, c_cur, bm->c_stage, h_stage16, int64_t(R), att_res16, nullptr, 0,
                              0.f, rows, s)))
      return rc;
    EpiStoreParams e3 = {};
    e3.alpha = 1.f; e3.bias = c->b_logit; e3.C = bm->logits; e3.ldc = V1;
    if ((rc = gemm_run(0, 0, 0, att_res16, R, c->w_logit16, R, rows, V1, R, 1, 0, e3, s))) return rc;
    const int step = t + 1;
    const int64_t hsz = int64_t(n_img) * T * bs;
    beam_step_kernel<<<n_img, BEAM_THREADS, 0, s>>>(
        bm->logits, V1, n_img, bs, step, T, bm->no_repeat, bm->hist_seq + (t & 1) * hsz,
        bm->hist_lp + (t & 1) * hsz, bm->hist_seq + ((t + 1) & 1) * hsz, bm->hist_lp + ((t + 1) & 1) * hsz,
        bm->beam_sum, bm->forced_parent, bm->forced_tok, bm->raw_parent, bm->raw_tok, bm->parent, bm->tok,
        bm->done_seq, bm->done_lp, bm->done_slot, bm->done_p_rec, bm->done_n);
    CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 8.0 * double(rows) * V1);
    if (step < T) {
      beam_reorder_kernel<<<rows, 128, 0, s>>>(bm->parent, bm->tok, h_stage16, bm->c_stage, c->embed, E, R, n_img,
                                              xh_nxt, c_nxt);
      CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 0.0);
    }
  }
  beam_finish_kernel<<<n_img, 32, 0, s>>>(bm->done_seq, bm->done_lp, bm->done_slot, bm->done_n, bm->beam_sum, bs, T,
                                         bm->done_p, bm->seq, bm->seq_logp);
  CC_LAUNCH_CHECK_K(PROF_SAMPLE, s, 0.0, 0.0);
  return CC_OK;
}

// scores[q, n] = sum_k Q[q, k] C[n, k] in fp32 FMAs, k ascending (the reference's np.dot of float32
// arrays; retrieval ranks must not depend on bf16 / tf32 rounding).  64 x 64 tile per CTA, 4 x 4
// outputs per thread, 16-deep k slices through shared memory.
__global__ void __launch_bounds__(256)
retrieval_scores_kernel(const float* __restrict__ Qm, const float* __restrict__ Cm, int nq, int nc, int K,
                        float* __restrict__ out, int64_t ld) {
  __shared__ float s_q[16][64 + 1];
  __shared__ float s_c[16][64 + 1];
  const int q0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      s_q[k][r] = (q0 + r < nq && k0 + k < K) ? Qm[int64_t(q0 + r) * K + k0 + k] : 0.f;
      s_c[k][r] = (c0 + r < nc && k0 + k < K) ? Cm[int64_t(c0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = s_q[k][ty * 4 + i]; b[i] = s_c[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (q < nq && c < nc) out[int64_t(q) * ld + c] = acc[i][j];
    }
}

// rank of the best correct candidate of every query: candidates scoring strictly higher
__global__ void __launch_bounds__(256)
retrieval_rank_kernel(const float* __restrict__ scores, int64_t ld, int n_cand, const int* __restrict__ first,
                      int count, int* __restrict__ ranks, int* __restrict__ top1) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  __shared__ int s_cnt[8];
  const int q = blockIdx.x;
  const float* row = scores + int64_t(q) * ld;
  const int f = first[q];
  float best = -INFINITY;                          // best correct score
  for (int j = 0; j < count; ++j)
    if (f + j >= 0 && f + j < n_cand) best = fmaxf(best, row[f + j]);
  int cnt = 0, ai = 0x7fffffff;
  float av = -INFINITY;
  for (int c = threadIdx.x; c < n_cand; c += 256) {
    const float v = row[c];
    cnt += v > best;
    if (v > av || (v == av && c < ai)) { av = v; ai = c; }
  }
  block_argmax(av, ai, s_v, s_i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_cnt[w];
    ranks[q] = t;
    top1[q] = ai;
  }
}

}  // namespace coopcap

extern "C" {

int coopcap_speaker_beam_fwd(const coopcap_speaker* ctx, const coopcap_beam* beam, coopcap_stream_t stream) {
  return coopcap::speaker_beam_fwd(ctx, beam, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_retrieval_scores(const float* queries, const float* cands, int n_query, int n_cand, int K,
                             float* scores, int64_t ld, coopcap_stream_t stream) {
  using namespace coopcap;
  CC_REQUIRE(queries && cands && scores && n_query > 0 && n_cand > 0 && K > 0 && ld >= n_cand,
             "retrieval_scores: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((n_cand + 63) / 64, (n_query + 63) / 64);
  CC_REQUIRE(grid.y <= 65535, "retrieval_scores: too many queries for one launch (%d)", n_query);
  retrieval_scores_kernel<<<grid, 256, 0, s>>>(queries, cands, n_query, n_cand, K, scores, ld);
  CC_LAUNCH_CHECK_K(PROF_HINGE, s, 2.0 * double(n_query) * n_cand * K, 0.0);
  return CC_OK;
}

int coopcap_retrieval_ranks(const float* scores, int64_t ld, int n_query, int n_cand, const int* first,
                            int count, int* ranks, int* top1, coopcap_stream_t stream) {
  using namespace coopcap;
  CC_REQUIRE(scores && first && ranks && top1 && n_query > 0 && n_cand > 0 && count > 0 && ld >= n_cand,
             "retrieval_ranks: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  retrieval_rank_kernel<<<n_query, 256, 0, s>>>(scores, ld, n_cand, first, count, ranks, top1);
  CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 4.0 * double(n_query) * n_cand);
  return CC_OK;
}

}  // extern "C"
