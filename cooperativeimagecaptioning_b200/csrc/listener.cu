// Listener (VSEFC): image encoder, GRU caption encoder over index captions, cosine score matrix,
// max-violation hinge loss -- forward and backward.  See include/coopcap.h for the buffer layout
// and the reference lines each piece replaces (models/VSEFCModel.py).
#include "../../include/coopcap.h"
#include "common.cuh"
#include "gemm.cuh"
#include "speaker_kernels.cuh"
#include "cell_step.cuh"

namespace coopcap {

using bf16 = __nv_bfloat16;

int cast_block(const float* src, int64_t rows, int cols, void* dst, int64_t ld_dst, cudaStream_t s);
int cast_blocks(int n, const float* const* src, const int64_t* rows, const int* cols, void* const* dst,
                const int64_t* ld_dst, cudaStream_t s);
int colsum_bf16(const void* src, int64_t rows, int cols, int64_t ld, float* out, cudaStream_t s);
int wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, float* C,
          int64_t ldc, cudaStream_t s);

// ------------------------------------------------------------------------------------------
// forward kernels
// ------------------------------------------------------------------------------------------
// emb16[s, b, :] = bf16(W_emb[tok[s, b], :])                           (VSEFCModel.py:102-106)
__global__ void gather_embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ w_emb,
                                    int E, bf16* __restrict__ emb16) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(w_emb + tok[row] * E);
  uint2* dst = reinterpret_cast<uint2*>(emb16 + row * E);
  for (int i = threadIdx.x; i < E / 4; i += blockDim.x) {
    const float4 v = __ldg(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    dst[i] = o;
  }
}

// one GRU step (torch gate order r, z, n), masked update for rows with t >= len
//   r = sig(gi_r + gh_r); z = sig(gi_z + gh_z); n = tanh(gi_n + r * gh_n); h' = (1-z) n + z h
__global__ void gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                               const float* __restrict__ h_prev, const int* __restrict__ len, int t,
                               float* __restrict__ gates, float* __restrict__ h_next,
                               bf16* __restrict__ h_next16, int B, int M) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = M / 4;
  if (idx >= B * per_row) return;
  const int b = idx / per_row, j = (idx % per_row) * 4;
  const float* gib = gi + int64_t(b) * 3 * M;
  const float* ghb = gh + int64_t(b) * 3 * M;
  const float4 ir = *reinterpret_cast<const float4*>(gib + j);
  const float4 iz = *reinterpret_cast<const float4*>(gib + M + j);
  const float4 in = *reinterpret_cast<const float4*>(gib + 2 * M + j);
  const float4 hr = *reinterpret_cast<const float4*>(ghb + j);
  const float4 hz = *reinterpret_cast<const float4*>(ghb + M + j);
  const float4 hn = *reinterpret_cast<const float4*>(ghb + 2 * M + j);
  const float4 hp = *reinterpret_cast<const float4*>(h_prev + int64_t(b) * M + j);
  const float a_ir[4] = {ir.x, ir.y, ir.z, ir.w}, a_iz[4] = {iz.x, iz.y, iz.z, iz.w};
  const float a_in[4] = {in.x, in.y, in.z, in.w}, a_hr[4] = {hr.x, hr.y, hr.z, hr.w};
  const float a_hz[4] = {hz.x, hz.y, hz.z, hz.w}, a_hn[4] = {hn.x, hn.y, hn.z, hn.w};
  const float a_hp[4] = {hp.x, hp.y, hp.z, hp.w};
  const bool active = t < len[b];
  float r[4], z[4], n[4], h[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    r[q] = 1.f / (1.f + expf(-(a_ir[q] + a_hr[q])));
    z[q] = 1.f / (1.f + expf(-(a_iz[q] + a_hz[q])));
    n[q] = tanhf(a_in[q] + r[q] * a_hn[q]);
    h[q] = active ? (1.f - z[q]) * n[q] + z[q] * a_hp[q] : a_hp[q];
  }
  float* g = gates + int64_t(b) * 4 * M;
  *reinterpret_cast<float4*>(g + j) = make_float4(r[0], r[1], r[2], r[3]);
  *reinterpret_cast<float4*>(g + M + j) = make_float4(z[0], z[1], z[2], z[3]);
  *reinterpret_cast<float4*>(g + 2 * M + j) = make_float4(n[0], n[1], n[2], n[3]);
  *reinterpret_cast<float4*>(g + 3 * M + j) = hn;
  *reinterpret_cast<float4*>(h_next + int64_t(b) * M + j) = make_float4(h[0], h[1], h[2], h[3]);
  __nv_bfloat162 a = __floats2bfloat162_rn(h[0], h[1]), c = __floats2bfloat162_rn(h[2], h[3]);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&c);
  *reinterpret_cast<uint2*>(h_next16 + int64_t(b) * M + j) = o;
}

// y = x / (||x||_2 + 1e-7), one CTA (256 threads) per row          (VSEFCModel.py:12-17)
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int M,
                                  int passthrough, int use_abs) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const float* xr = x + int64_t(blockIdx.x) * M;
  float* yr = y + int64_t(blockIdx.x) * M;
  float ss = 0.f;
  for (int i = threadIdx.x; i < M; i += 256) ss += xr[i] * xr[i];
  ss = block_sum_256(ss, red);
  const float inv = passthrough ? 1.f : 1.f / (sqrtf(ss) + 1e-7f);
  // use_abs: |.| after the normalisation (VSEFCModel.py:50-52,137-139)
  for (int i = threadIdx.x; i < M; i += 256) yr[i] = use_abs ? fabsf(xr[i] * inv) : xr[i] * inv;
}

// dx = dy / (n + eps) - x (dy . x) / (n (n + eps)^2); optional fp32 and bf16 outputs
__global__ void l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                  float* __restrict__ dx, bf16* __restrict__ dx16, int M,
                                  int passthrough, int use_abs) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  const float* xr = x + int64_t(blockIdx.x) * M;
  const float* dr = dy + int64_t(blockIdx.x) * M;
  // use_abs: the upstream gradient first passes |.|, i.e. takes the sign of the normalised value
  auto dyv = [&](int i) { return (use_abs && xr[i] < 0.f) ? -dr[i] : dr[i]; };
  float ss = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < M; i += 256) {
    ss += xr[i] * xr[i];
    dot += xr[i] * dyv(i);
  }
  ss = block_sum_256(ss, red);
  dot = block_sum_256(dot, red);
  const float n = sqrtf(ss), ne = n + 1e-7f;
  const float a = passthrough ? 1.f : 1.f / ne;
  const float c = (passthrough || n == 0.f) ? 0.f : dot / (n * ne * ne);
  for (int i = threadIdx.x; i < M; i += 256) {
    const float v = dyv(i) * a - xr[i] * c;
    if (dx) dx[int64_t(blockIdx.x) * M + i] = v;
    if (dx16) dx16[int64_t(blockIdx.x) * M + i] = __float2bfloat16_rn(v);
  }
}

// max-violation hinge terms                                         (VSEFCModel.py:167-193)
//   cost_s[i]  = relu(margin + max_{j != i} S_ij - S_ii)   (caption retrieval, row max)
//   cost_im[j] = relu(margin + max_{i != j} S_ij - S_jj)   (image retrieval, column max)
__global__ void hinge_rows_kernel(const float* __restrict__ S, int B, float margin,
                                  float* __restrict__ cost_s, int* __restrict__ arg_s) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sv[8];
  __shared__ int si[8];
  const int i = blockIdx.x;
  const float* row = S + int64_t(i) * B;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < B; j += 256) {
    if (j == i) continue;
    const float v = row[j];
    if (v > bv) { bv = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
    const float c = (bi == 0x7fffffff) ? 0.f : fmaxf(margin + bv - row[i], 0.f);
    cost_s[i] = c;
    arg_s[i] = (bi == 0x7fffffff) ? i : bi;
  }
}

constexpr int HC_GROUPS = 32;
__global__ void hinge_cols_kernel(const float* __restrict__ S, int B, float margin,
                                  float* __restrict__ cost_im, int* __restrict__ arg_im) {
  pdl_launch_dependents();
  pdl_wait();
  // HC_GROUPS row groups per column block: 32 CTAs x 1024 threads walk the B x B matrix (8 groups
  // left each thread 128 dependent compares on 32 of 148 SMs: 23 us for 4 MB)
  __shared__ float sv[HC_GROUPS][33];
  __shared__ int si[HC_GROUPS][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  if (j < B) {
    for (int i = threadIdx.y; i < B; i += HC_GROUPS) {
      if (i == j) continue;
      const float v = S[int64_t(i) * B + j];
      if (v > bv) { bv = v; bi = i; }
    }
  }
  sv[threadIdx.y][threadIdx.x] = bv;
  si[threadIdx.y][threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.y == 0 && j < B) {
    for (int w = 1; w < HC_GROUPS; ++w) {
      const float ov = sv[w][threadIdx.x];
      const int oi = si[w][threadIdx.x];
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    const float c = (bi == 0x7fffffff) ? 0.f : fmaxf(margin + bv - S[int64_t(j) * B + j], 0.f);
    cost_im[j] = c;
    arg_im[j] = (bi == 0x7fffffff) ? j : bi;
  }
}

// vse_max_violation = 0 (VSEFCModel.py:190-193 else-branch): mean over the row / column of the
// hinge matrix with its diagonal zeroed, i.e. (1/B) sum over the negatives
__global__ void hinge_sum_rows_kernel(const float* __restrict__ S, int B, float margin,
                                      float* __restrict__ cost_s) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float* row = S + int64_t(i) * B;
  const float d = row[i];
  float acc = 0.f;
  for (int j = threadIdx.x; j < B; j += 256)
    if (j != i) acc += fmaxf(margin + row[j] - d, 0.f);
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) cost_s[i] = acc / float(B);
}
__global__ void hinge_sum_cols_kernel(const float* __restrict__ S, int B, float margin,
                                      float* __restrict__ cost_im) {
  __shared__ float sv[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (j < B) {
    const float d = S[int64_t(j) * B + j];
    for (int i = threadIdx.y; i < B; i += 8)
      if (i != j) acc += fmaxf(margin + S[int64_t(i) * B + j] - d, 0.f);
  }
  sv[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < B) {
    for (int w = 1; w < 8; ++w) acc += sv[w][threadIdx.x];
    cost_im[j] = acc / float(B);
  }
}
// its backward: dS[i,j] (i != j) = g_i [m + S_ij - S_ii > 0] / B + g_j [m + S_ij - S_jj > 0] / B,
// dS[i,i] = -(sum of row i's first terms) - (sum of column i's second terms); one CTA per row i
// writes the off-diagonal entries and the row part of the diagonal, a second pass adds the column part
__global__ void hinge_sum_ds_kernel(const float* __restrict__ S, int B, float margin,
                                    const float* __restrict__ g_loss, const float* __restrict__ g_rows,
                                    int only, float* __restrict__ dS) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const float gi = g_loss ? g_loss[0] : g_rows[i];
  const float dii = S[int64_t(i) * B + i];
  const float invB = 1.f / float(B);
  float rowpart = 0.f;
  for (int j = threadIdx.x; j < B; j += 256) {
    if (j == i) continue;
    const float sij = S[int64_t(i) * B + j];
    const float gj = g_loss ? g_loss[0] : g_rows[j];
    float v = 0.f;
    if (only != 1 && margin + sij - dii > 0.f) { v += gi * invB; rowpart += gi * invB; }
    if (only != 2 && margin + sij - S[int64_t(j) * B + j] > 0.f) v += gj * invB;
    dS[int64_t(i) * B + j] = v;
  }
  rowpart = block_sum_256(rowpart, red);
  if (threadIdx.x == 0) dS[int64_t(i) * B + i] = -rowpart;
}
__global__ void hinge_sum_ds_diag_kernel(const float* __restrict__ S, int B, float margin,
                                         const float* __restrict__ g_loss,
                                         const float* __restrict__ g_rows, int only,
                                         float* __restrict__ dS) {
  __shared__ float red[8];
  const int j = blockIdx.x;
  float colpart = 0.f;
  if (only != 2) {
    const float gj = g_loss ? g_loss[0] : g_rows[j];
    const float djj = S[int64_t(j) * B + j];
    for (int i = threadIdx.x; i < B; i += 256)
      if (i != j && margin + S[int64_t(i) * B + j] - djj > 0.f) colpart += gj / float(B);
  }
  colpart = block_sum_256(colpart, red);
  if (threadIdx.x == 0) dS[int64_t(j) * B + j] -= colpart;
}

// vse_pool_type mean / max over the valid steps (VSEFCModel.py:115-126); h32[t + 1] is the state
// after step t
__global__ void pool_kernel(const float* __restrict__ h32, const int* __restrict__ len, int S, int B,
                            int M, int pool_type, float* __restrict__ cap_pre,
                            int* __restrict__ pool_arg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * M) return;
  const int b = idx / M;
  const int n = min(len[b], S);
  if (pool_type == 1) {
    float acc = 0.f;
    for (int t = 0; t < n; ++t) acc += h32[int64_t(t + 1) * B * M + idx];
    cap_pre[idx] = acc / float(n);
  } else {
    float bv = -1e10f;     // the reference's fill value for masked steps
    int bt = 0;
    for (int t = 0; t < n; ++t) {
      const float v = h32[int64_t(t + 1) * B * M + idx];
      if (v > bv) { bv = v; bt = t; }
    }
    cap_pre[idx] = bv;
    pool_arg[idx] = bt;
  }
}

// loss_rows = cost_s (+) cost_im according to only_one_retrieval; loss = sum (deterministic order)
__global__ void hinge_finish_kernel(const float* __restrict__ cost_s, const float* __restrict__ cost_im,
                                    int B, int only, float* __restrict__ loss_rows,
                                    float* __restrict__ loss) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float v = (only == 1 ? 0.f : cost_s[i]) + (only == 2 ? 0.f : cost_im[i]);
    loss_rows[i] = v;
    acc += v;
  }
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) loss[0] = acc;
}

// ------------------------------------------------------------------------------------------
// backward kernels
// ------------------------------------------------------------------------------------------
// sparse backward of the max-violation hinge: one off-diagonal entry + the diagonal per row and
// per column.  d_im / d_cap are zeroed by the caller; contributions are atomically added.
__global__ void hinge_bwd_kernel(const float* __restrict__ im, const float* __restrict__ cap,
                                 const float* __restrict__ cost_s, const float* __restrict__ cost_im,
                                 const int* __restrict__ arg_s, const int* __restrict__ arg_im,
                                 const float* __restrict__ g_loss, const float* __restrict__ g_rows,
                                 int only, int M, float* __restrict__ d_im, float* __restrict__ d_cap) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x;
  const float g = g_loss ? g_loss[0] : g_rows[i];
  const bool do_s = (only != 1) && cost_s[i] > 0.f && g != 0.f;
  const bool do_i = (only != 2) && cost_im[i] > 0.f && g != 0.f;
  if (do_s) {
    const int j = arg_s[i];   // d cost_s[i] = g * (dS[i,j] - dS[i,i]),  S[i,j] = im[i] . cap[j]
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
      const float imi = im[int64_t(i) * M + k];
      atomicAdd(d_im + int64_t(i) * M + k, g * (cap[int64_t(j) * M + k] - cap[int64_t(i) * M + k]));
      atomicAdd(d_cap + int64_t(j) * M + k, g * imi);
      atomicAdd(d_cap + int64_t(i) * M + k, -g * imi);
    }
  }
  if (do_i) {
    const int r = arg_im[i];  // column i: d cost_im[i] = g * (dS[r,i] - dS[i,i])
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
      const float ci = cap[int64_t(i) * M + k];
      atomicAdd(d_cap + int64_t(i) * M + k, g * (im[int64_t(r) * M + k] - im[int64_t(i) * M + k]));
      atomicAdd(d_im + int64_t(r) * M + k, g * ci);
      atomicAdd(d_im + int64_t(i) * M + k, -g * ci);
    }
  }
}

// one reverse GRU step.  dh (in/out): on entry d(loss)/d(h_t); on exit the direct part of
// d(loss)/d(h_{t-1}) (the recurrent part, d_gh . W_hh, is accumulated by the following GEMM).
__global__ void gru_bwd_kernel(float* __restrict__ dh, const float* __restrict__ gates,
                               const float* __restrict__ h_prev, const int* __restrict__ len, int t,
                               bf16* __restrict__ d_gi16, bf16* __restrict__ d_gh16, int B, int M,
                               int pool_type, const float* __restrict__ d_pool,
                               const int* __restrict__ pool_arg) {
  pdl_launch_dependents();
  pdl_wait();
  // 4 hidden units per thread: 16-byte loads, 8-byte bf16 stores
  const int idx4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = M / 4;
  if (idx4 >= B * per_row) return;
  const int b = idx4 / per_row, j = (idx4 % per_row) * 4;
  const int64_t idx = int64_t(b) * M + j;
  const bool active = t < len[b];
  bf16* gi = d_gi16 + int64_t(b) * 3 * M;
  bf16* gh = d_gh16 + int64_t(b) * 3 * M;
  if (!active) {
    store_bf16x4(gi + j, 0.f, 0.f, 0.f, 0.f); store_bf16x4(gi + M + j, 0.f, 0.f, 0.f, 0.f);
    store_bf16x4(gi + 2 * M + j, 0.f, 0.f, 0.f, 0.f);
    store_bf16x4(gh + j, 0.f, 0.f, 0.f, 0.f); store_bf16x4(gh + M + j, 0.f, 0.f, 0.f, 0.f);
    store_bf16x4(gh + 2 * M + j, 0.f, 0.f, 0.f, 0.f);
    return;  // dh passes through unchanged
  }
  auto ld4 = [](const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  };
  const float* g = gates + int64_t(b) * 4 * M;
  float r[4], z[4], n[4], ghn[4], hp[4], d[4];
  ld4(g + j, r); ld4(g + M + j, z); ld4(g + 2 * M + j, n); ld4(g + 3 * M + j, ghn);
  ld4(h_prev + idx, hp); ld4(dh + idx, d);
  // pooled variants: h_t also feeds the pool directly
  if (pool_type == 1) {
    float dp[4];
    ld4(d_pool + idx, dp);
    const float inv = 1.f / float(len[b]);
#pragma unroll
    for (int q = 0; q < 4; ++q) d[q] += dp[q] * inv;
  } else if (pool_type == 2) {
    float dp[4];
    ld4(d_pool + idx, dp);
    const int4 pa = *reinterpret_cast<const int4*>(pool_arg + idx);
    const int a4[4] = {pa.x, pa.y, pa.z, pa.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (a4[q] == t) d[q] += dp[q];
  }
  float o_r[4], o_z[4], o_n[4], o_nr[4], o_dh[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float dn = d[q] * (1.f - z[q]);
    const float dz = d[q] * (hp[q] - n[q]);
    const float dnp = dn * (1.f - n[q] * n[q]);
    o_z[q] = dz * z[q] * (1.f - z[q]);
    o_r[q] = dnp * ghn[q] * r[q] * (1.f - r[q]);
    o_n[q] = dnp;
    o_nr[q] = dnp * r[q];
    o_dh[q] = d[q] * z[q];
  }
  store_bf16x4(gi + j, o_r[0], o_r[1], o_r[2], o_r[3]);
  store_bf16x4(gi + M + j, o_z[0], o_z[1], o_z[2], o_z[3]);
  store_bf16x4(gi + 2 * M + j, o_n[0], o_n[1], o_n[2], o_n[3]);
  store_bf16x4(gh + j, o_r[0], o_r[1], o_r[2], o_r[3]);
  store_bf16x4(gh + M + j, o_z[0], o_z[1], o_z[2], o_z[3]);
  store_bf16x4(gh + 2 * M + j, o_nr[0], o_nr[1], o_nr[2], o_nr[3]);
  *reinterpret_cast<float4*>(dh + idx) = make_float4(o_dh[0], o_dh[1], o_dh[2], o_dh[3]);
}

// g_w_emb[tok[s,b], :] += demb[s, b, :] for s < len[b]            (sparse embedding wgrad)
__global__ void embed_scatter_kernel(const int64_t* __restrict__ tok, const int* __restrict__ len,
                                     const bf16* __restrict__ demb16, int B, int E,
                                     float* __restrict__ g_w_emb) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t row = blockIdx.x;
  const int s = int(row / B), b = int(row % B);
  if (s >= len[b]) return;
  float* dst = g_w_emb + tok[row] * E;
  if ((E & 3) == 0) {
    // four columns per 16-byte vector atomic (red.global.add.v4.f32)
    for (int i = threadIdx.x * 4; i < E; i += blockDim.x * 4) {
      const uint2 r = *reinterpret_cast<const uint2*>(demb16 + row * E + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
      const float2 a = __bfloat1622float2(h[0]), b2 = __bfloat1622float2(h[1]);
      atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(a.x, a.y, b2.x, b2.y));
    }
    return;
  }
  for (int i = threadIdx.x; i < E; i += blockDim.x)
    atomicAdd(dst + i, __bfloat162float(demb16[row * E + i]));
}

// ------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------
static int check_dims(const coopcap_listener* c) {
  CC_REQUIRE(c != nullptr, "listener: null context");
  CC_REQUIRE(c->B > 0 && c->S > 0, "listener: empty batch B=%d S=%d", c->B, c->S);
  CC_REQUIRE(c->F % 8 == 0 && c->M % 8 == 0 && c->E % 8 == 0, "listener: F,M,E must be multiples of 8");
  CC_REQUIRE(c->pool_type >= 0 && c->pool_type <= 2, "listener: pool_type %d", c->pool_type);
  CC_REQUIRE(c->pool_type == 0 || (c->cap_pre && (c->pool_type == 1 || c->pool_arg)),
             "listener: pool mean / max need cap_pre (and pool_arg)");
  return CC_OK;
}

int listener_fwd(const coopcap_listener* c, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  const int B = c->B, S = c->S, M = c->M, E = c->E, F = c->F;
  bf16* h16 = reinterpret_cast<bf16*>(c->h16);
  // image branch                                                      (VSEFCModel.py:40-54)
  if ((rc = cast_block(c->fc_feats, B, F, c->fc16, F, s))) return rc;
  {
    EpiStoreParams e = {};
    e.alpha = 1.f; e.bias = c->b_img; e.C = c->img_pre; e.ldc = M;
    if ((rc = gemm_run(0, 0, 0, c->fc16, F, c->w_img16, F, B, M, F, 1, 0, e, s))) return rc;
  }
  CC_CHECK_CUDA(launch_pdl(l2norm_fwd_kernel, dim3(B), dim3(256), size_t(0), s, c->img_pre, c->im, M, c->no_imgnorm, c->use_abs));
  CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  // caption branch                                                    (VSEFCModel.py:83-140)
  if (!c->emb_given) {
    CC_CHECK_CUDA(launch_pdl(gather_embed_kernel, dim3(S * B), dim3(128), size_t(0), s, c->tok, c->w_emb, E, reinterpret_cast<bf16*>(c->emb16)));
    CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 0.0);
  }
  {
    EpiStoreParams e = {};
    e.alpha = 1.f; e.bias = c->b_ih; e.C = c->gi_all; e.ldc = 3 * M;
    if ((rc = gemm_run(0, 0, 0, c->emb16, E, c->w_ih16, E, S * B, 3 * M, E, 1, 0, e, s))) return rc;
  }
  CC_CHECK_CUDA(cudaMemsetAsync(c->h32, 0, sizeof(float) * B * M, s));
  CC_CHECK_CUDA(cudaMemsetAsync(h16, 0, sizeof(bf16) * B * M, s));
  for (int t = 0; t < S; ++t) {
    if (cell_step_ok<GruCell>(M, M)) {
      // gh GEMM with the GRU update as its epilogue (csrc/cell_step.cuh); c->gh is not written
      GruStepParams gp = {};
      gp.b_hh = c->b_hh; gp.gi = c->gi_all + int64_t(t) * B * 3 * M; gp.h_prev = c->h32 + int64_t(t) * B * M;
      gp.len = c->len; gp.gates = c->gates + int64_t(t) * B * 4 * M;
      gp.h_next = c->h32 + int64_t(t + 1) * B * M; gp.h_next16 = h16 + int64_t(t + 1) * B * M; gp.t = t;
      if ((rc = launch_cell_step<GruCell>(h16 + int64_t(t) * B * M, M, c->w_hh16, B, M, M, gp, s))) return rc;
      continue;
    }
    EpiStoreParams e = {};
    e.alpha = 1.f; e.bias = c->b_hh; e.C = c->gh; e.ldc = 3 * M;
    if ((rc = gemm_run(0, 0, 0, h16 + int64_t(t) * B * M, M, c->w_hh16, M, B, 3 * M, M, 1, 0, e, s)))
      return rc;
    const int n = B * (M / 4);
    CC_CHECK_CUDA(launch_pdl(gru_fwd_kernel, dim3((n + 255) / 256), dim3(256), 0, s,
                             c->gi_all + int64_t(t) * B * 3 * M, c->gh, c->h32 + int64_t(t) * B * M,
                             c->len, t, c->gates + int64_t(t) * B * 4 * M,
                             c->h32 + int64_t(t + 1) * B * M, h16 + int64_t(t + 1) * B * M, B, M));
    CC_LAUNCH_CHECK_K(PROF_GRU, s, 0.0, 0.0);
  }
  const float* pooled = c->h32 + int64_t(S) * B * M;     // 'last': the masked update carries it
  if (c->pool_type != 0) {
    pool_kernel<<<(B * M + 255) / 256, 256, 0, s>>>(c->h32, c->len, S, B, M, c->pool_type, c->cap_pre,
                                                    c->pool_arg);
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
    pooled = c->cap_pre;
  }
  CC_CHECK_CUDA(launch_pdl(l2norm_fwd_kernel, dim3(B), dim3(256), size_t(0), s, pooled, c->cap, M, 0, c->use_abs));
  CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  // scores (tf32 operands: the hinge compares score differences against a 0.2 margin)
  {
    EpiStoreParams e = {};
    e.alpha = 1.f; e.C = c->scores; e.ldc = B;
    if ((rc = gemm_run(1, 0, 0, c->im, M, c->cap, M, B, B, M, 1, 0, e, s))) return rc;
  }
  if (c->sum_violation) {
    hinge_sum_rows_kernel<<<B, 256, 0, s>>>(c->scores, B, c->margin, c->cost_s);
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
    hinge_sum_cols_kernel<<<(B + 31) / 32, dim3(32, 8), 0, s>>>(c->scores, B, c->margin, c->cost_im);
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  } else {
    CC_CHECK_CUDA(launch_pdl(hinge_rows_kernel, dim3(B), dim3(256), size_t(0), s, c->scores, B, c->margin, c->cost_s, c->arg_s));
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
    CC_CHECK_CUDA(launch_pdl(hinge_cols_kernel, dim3((B + 31) / 32), dim3(32, HC_GROUPS), size_t(0), s, c->scores, B, c->margin, c->cost_im,
                                                             c->arg_im));
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  }
  CC_CHECK_CUDA(launch_pdl(hinge_finish_kernel, dim3(1), dim3(256), size_t(0), s, c->cost_s, c->cost_im, B, c->only_one_retrieval,
                                        c->loss_rows, c->loss));
  CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  return CC_OK;
}

int listener_bwd(const coopcap_listener* c, const coopcap_listener_grads* g, cudaStream_t s) {
  int rc = check_dims(c);
  if (rc) return rc;
  CC_REQUIRE(g != nullptr, "listener_bwd: null grads");
  CC_REQUIRE((g->g_loss != nullptr) != (g->g_rows != nullptr),
             "listener_bwd: exactly one of g_loss / g_rows must be given");
  const int B = c->B, S = c->S, M = c->M, E = c->E, F = c->F;
  bf16* h16 = reinterpret_cast<bf16*>(c->h16);
  bf16* d_gi16 = reinterpret_cast<bf16*>(g->d_gi16);
  bf16* d_gh16 = reinterpret_cast<bf16*>(g->d_gh16);
  if (c->sum_violation) {
    // dense d(scores), then d_im = dS . cap and d_cap = dS^T . im (fp32 SIMT: B x B x M, non-default)
    CC_REQUIRE(g->d_scores != nullptr, "listener_bwd: the sum-violation hinge needs d_scores");
    hinge_sum_ds_kernel<<<B, 256, 0, s>>>(c->scores, B, c->margin, g->g_loss, g->g_rows,
                                          c->only_one_retrieval, g->d_scores);
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
    hinge_sum_ds_diag_kernel<<<B, 256, 0, s>>>(c->scores, B, c->margin, g->g_loss, g->g_rows,
                                               c->only_one_retrieval, g->d_scores);
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
    coopcap_gemm_args a = {};
    a.kind = 1; a.backend = 1; a.alpha = 1.f;
    a.M = B; a.N = M; a.K = B;
    a.A = g->d_scores; a.lda = B; a.a_major = 0; a.B = c->cap; a.ldb = M; a.b_major = 1;
    a.C = g->d_im; a.ldc = M;
    if ((rc = gemm_store(&a, s))) return rc;
    a.a_major = 1; a.B = c->im; a.C = g->d_cap;
    if ((rc = gemm_store(&a, s))) return rc;
  } else {
    CC_CHECK_CUDA(cudaMemsetAsync(g->d_im, 0, sizeof(float) * B * M, s));
    CC_CHECK_CUDA(cudaMemsetAsync(g->d_cap, 0, sizeof(float) * B * M, s));
    CC_CHECK_CUDA(launch_pdl(hinge_bwd_kernel, dim3(B), dim3(256), size_t(0), s, c->im, c->cap, c->cost_s, c->cost_im, c->arg_s, c->arg_im,
                                       g->g_loss, g->g_rows, c->only_one_retrieval, M, g->d_im,
                                       g->d_cap));
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  }
  if (g->need_param_grads) {
    CC_CHECK_CUDA(launch_pdl(l2norm_bwd_kernel, dim3(B), dim3(256), size_t(0), s, c->img_pre, g->d_im, nullptr,
                                        reinterpret_cast<bf16*>(g->d_img_pre16), M, c->no_imgnorm,
                                        c->use_abs));
    CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
    if ((rc = wgrad(g->d_img_pre16, M, c->fc16, F, M, F, B, g->g_w_img, F, s))) return rc;
    if ((rc = colsum_bf16(g->d_img_pre16, B, M, M, g->g_b_img, s))) return rc;
  }
  if (c->pool_type == 0) {
    CC_CHECK_CUDA(launch_pdl(l2norm_bwd_kernel, dim3(B), dim3(256), size_t(0), s, c->h32 + int64_t(S) * B * M, g->d_cap, g->dh, nullptr, M, 0,
                                        c->use_abs));
  } else {
    // the pooled state feeds every valid step (mean) / its arg-max step (max): gru_bwd_kernel adds it
    CC_REQUIRE(g->d_pool != nullptr, "listener_bwd: pool mean / max need d_pool");
    CC_CHECK_CUDA(launch_pdl(l2norm_bwd_kernel, dim3(B), dim3(256), size_t(0), s, c->cap_pre, g->d_cap, g->d_pool, nullptr, M, 0, c->use_abs));
    CC_CHECK_CUDA(cudaMemsetAsync(g->dh, 0, sizeof(float) * B * M, s));
  }
  CC_LAUNCH_CHECK_K(PROF_HINGE, s, 0.0, 0.0);
  for (int t = S - 1; t >= 0; --t) {
    const int n = B * (M / 4);
    CC_CHECK_CUDA(launch_pdl(gru_bwd_kernel, dim3((n + 255) / 256), dim3(256), 0, s, g->dh,
                             c->gates + int64_t(t) * B * 4 * M, c->h32 + int64_t(t) * B * M, c->len, t,
                             d_gi16 + int64_t(t) * B * 3 * M, d_gh16 + int64_t(t) * B * 3 * M, B, M,
                             c->pool_type, g->d_pool, c->pool_arg));
    CC_LAUNCH_CHECK_K(PROF_GRU, s, 0.0, 0.0);
    if (t > 0) {
      // dh += d_gh . W_hh        ([B,3M] x [3M,M]; W_hh stored [K, N])
      // few output tiles, deep K: split along K, both halves reduce-add onto dh
      const int split = (int64_t(B) * M <= int64_t(128) * 128 * 74 && 3 * M >= 2048) ? 2 : 1;
      EpiStoreParams e = {};
      e.alpha = 1.f; e.C = g->dh; e.ldc = M; e.mode = split > 1 ? 2 : 1;
      if ((rc = gemm_run(0, 0, 1, d_gh16 + int64_t(t) * B * 3 * M, 3 * M, c->w_hh16, M, B, M, 3 * M,
                         split, split > 1 ? 128 : 0, e, s)))
        return rc;
    }
  }
  // d(word embedding input) = d_gi . W_ih   ([S*B,3M] x [3M,E])
  {
    EpiStoreParams e = {};
    e.alpha = 1.f; e.C16 = reinterpret_cast<bf16*>(g->demb16); e.ldc16 = E;
    if ((rc = gemm_run(0, 0, 1, d_gi16, 3 * M, c->w_ih16, E, S * B, E, 3 * M, 1, 0, e, s))) return rc;
  }
  if (g->need_param_grads) {
    const int K = S * B;
    if ((rc = wgrad(d_gi16, 3 * M, c->emb16, E, 3 * M, E, K, g->g_w_ih, E, s))) return rc;
    if ((rc = wgrad(d_gh16, 3 * M, h16, M, 3 * M, M, K, g->g_w_hh, M, s))) return rc;
    if ((rc = colsum_bf16(d_gi16, K, 3 * M, 3 * M, g->g_b_ih, s))) return rc;
    if ((rc = colsum_bf16(d_gh16, K, 3 * M, 3 * M, g->g_b_hh, s))) return rc;
    if (!c->emb_given) {   // dense captions: coopcap_caption_embed_dense_bwd forms this gradient
      CC_CHECK_CUDA(launch_pdl(embed_scatter_kernel, dim3(S * B), dim3(128), size_t(0), s, c->tok, c->len,
                                                 reinterpret_cast<const bf16*>(g->demb16), B, E,
                                                 g->g_w_emb));
      CC_LAUNCH_CHECK_K(PROF_REDUCE, s, 0.0, 0.0);
    }
  }
  return CC_OK;
}

}  // namespace coopcap

extern "C" {

int coopcap_listener_pack_weights(const coopcap_listener_pack* p, coopcap_stream_t stream) {
  using namespace coopcap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CC_REQUIRE(p != nullptr, "listener_pack: null");
  const float* src[4] = {p->w_img, p->w_ih, p->w_hh, p->w_emb};
  const int64_t rows[4] = {p->M, 3 * p->M, 3 * p->M, p->V2};
  const int cols[4] = {p->F, p->E, p->M, p->E};
  void* dst[4] = {p->w_img16, p->w_ih16, p->w_hh16, p->w_emb16};
  const int64_t ld[4] = {p->F, p->E, p->M, p->E};
  return cast_blocks(p->w_emb16 ? 4 : 3, src, rows, cols, dst, ld, s);
}

// emb16[row, :] = bf16(w_emb[id, :]) for `rows` rows (the prepended BOS position)
__global__ void fill_embed_kernel(const float* __restrict__ w_row, int E,
                                  __nv_bfloat16* __restrict__ emb16) {
  __nv_bfloat16* dst = emb16 + int64_t(blockIdx.x) * E;
  for (int i = threadIdx.x; i < E; i += blockDim.x) dst[i] = __float2bfloat16_rn(w_row[i]);
}

int coopcap_caption_embed_dense(const void* soft16, const float* w_emb, const void* w_emb16, int n,
                                int B, int V1, int E, int64_t bos_id, void* emb16,
                                coopcap_stream_t stream) {
  using namespace coopcap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CC_REQUIRE(soft16 && w_emb && w_emb16 && emb16 && n >= 1 && B > 0 && V1 % 8 == 0 && E % 8 == 0,
             "caption_embed_dense: bad arguments");
  __nv_bfloat16* e16 = reinterpret_cast<__nv_bfloat16*>(emb16);
  fill_embed_kernel<<<B, 128, 0, s>>>(w_emb + bos_id * E, E, e16);
  CC_LAUNCH_CHECK_K(PROF_PACK, s, 0.0, 0.0);
  EpiStoreParams e = {};
  e.alpha = 1.f; e.C16 = e16 + int64_t(B) * E; e.ldc16 = E;
  return gemm_run(0, 0, 1, soft16, V1, w_emb16, E, n * B, E, V1, 1, 0, e, s);
}

int coopcap_caption_embed_dense_bwd(const void* soft16, const void* demb16, int n, int B, int V1,
                                    int E, int64_t bos_id, float* g_w_emb, coopcap_stream_t stream) {
  using namespace coopcap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CC_REQUIRE(soft16 && demb16 && g_w_emb && n >= 1 && B > 0, "caption_embed_dense_bwd: bad arguments");
  const __nv_bfloat16* d16 = reinterpret_cast<const __nv_bfloat16*>(demb16);
  int rc = wgrad(soft16, V1, d16 + int64_t(B) * E, E, V1, E, n * B, g_w_emb, E, s);
  if (rc) return rc;
  return colsum_bf16(d16, B, E, E, g_w_emb + bos_id * E, s);
}

int coopcap_listener_fwd(const coopcap_listener* ctx, coopcap_stream_t stream) {
  return coopcap::listener_fwd(ctx, reinterpret_cast<cudaStream_t>(stream));
}

int coopcap_listener_bwd(const coopcap_listener* ctx, const coopcap_listener_grads* gr,
                         coopcap_stream_t stream) {
  return coopcap::listener_bwd(ctx, gr, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
