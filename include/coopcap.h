/*
 * coopcap.h -- C ABI of libcoopcap.so, the sm_100a (B200) kernel library behind the joint
 * speaker-listener training step of CooperativeImageCaptioning.
 *
 * The reference has no FFI of its own (it is pure PyTorch); its "operator interface" for this
 * path is the set of torch calls made by models/AttModel.py, models/VSEFCModel.py,
 * models/gumbel.py, models/multinomial.py, misc/utils.py and optimizer.py.  Each entry point
 * below names the reference code (file:line under the reference root) whose arithmetic it
 * replaces.  The Python host side (cooperativeimagecaptioning_b200/*.py) binds these with ctypes
 * and wraps them in torch.autograd.Functions; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, scalars, an opaque stream handle (cudaStream_t).
 *   - every function returns 0 on success or a negative COOPCAP_ERR_* code; it never throws and
 *     never synchronises the device.  coopcap_last_error() returns a thread-local message.
 *   - all buffers (inputs, outputs, workspaces) are caller-allocated device memory; the library
 *     keeps no global device state and is re-entrant per stream.
 *   - row-major everywhere; "ld" arguments are leading dimensions in elements.
 *   - bf16 buffers are passed as void* (uint16 storage).
 */
#ifndef COOPCAP_H_
#define COOPCAP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COOPCAP_VERSION 100

#define COOPCAP_OK 0
#define COOPCAP_ERR_CUDA (-1)
#define COOPCAP_ERR_ARG (-2)
#define COOPCAP_ERR_DRIVER (-3)
#define COOPCAP_ERR_UNSUPPORTED (-4)

typedef void* coopcap_stream_t; /* cudaStream_t */

/* ---- runtime ------------------------------------------------------------------------------ */
int coopcap_version(void);
const char* coopcap_last_error(void);
int coopcap_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- dense contraction engine ---------------------------------------------------------------
 * C[M,N] = alpha * sum_k A[m,k] B[n,k] (+ bias[n]) (relu) (* row_scale[m])
 * Replaces every nn.Linear / mm / bmm on the path: AttModel.py:82-88 (att_embed, logit, ctx2att),
 * AttModel.py:470,503-505,514,522 (h2att, a2c, i2h, h2h), VSEFCModel.py:28,44 (img_enc.fc),
 * VSEFCModel.py:74-76 (GRU projections), VSEFCModel.py:104 (one-hot @ embed), VSEFCModel.py:143-146
 * (cosine_sim) and their autograd dgrad / wgrad contractions.
 * kind 0: bf16 operands; kind 1: fp32 operands consumed as tf32.  fp32 accumulation (TMEM).
 * a_major/b_major 0: operand stored [rows, K] (K contiguous); 1: stored [K, rows].
 * backend 0: tcgen05/TMA kernel; backend 1: plain SIMT kernel kept only to cross-check backend 0
 * in tests (never used by the product path).
 */
typedef struct coopcap_gemm_args {
  int kind;
  int a_major, b_major;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  int M, N, K;
  float alpha;
  const float* bias;      /* [N] or NULL */
  const float* row_scale; /* [M] or NULL */
  int relu;
  int mode;    /* 0 store, 1 C += result, 2 atomicAdd into C (required when split_k > 1) */
  float* C;    /* fp32 [M, ldc] or NULL */
  int64_t ldc;
  void* C16;   /* bf16 [M, ldc16] or NULL */
  int64_t ldc16;
  void* Ct16;  /* bf16 transposed [N, ldct] or NULL */
  int64_t ldct;
  int split_k; /* >= 1 */
  int tile_n;  /* 0 = auto, else 64 / 128 / 256 */
  int backend; /* 0 tcgen05, 1 SIMT cross-check */
} coopcap_gemm_args;

int coopcap_gemm(const coopcap_gemm_args* args, coopcap_stream_t stream);

/* fp32 -> bf16 copy (weights, activations); optional transposed copy dst_t [cols, rows]. */
int coopcap_cast_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst,
                      int64_t ld_dst, void* dst_t, int64_t ld_dst_t, coopcap_stream_t stream);


/* ---- speaker (Att2in2) -----------------------------------------------------------------------
 * One context struct describes a whole speaker pass over B rows: dimensions, inputs, parameters
 * (fp32 masters + packed bf16 operand copies), RNG configuration, and every activation that the
 * backward pass needs.  All buffers are caller-allocated; the library only launches kernels.
 *
 * Layout in HBM (row-major, sizes in elements):
 *   att16      bf16 [B*L, D]            cast of att_feats (operand of att_embed fwd and wgrad)
 *   att_e16    bf16 [B*L, R]            dropout(relu(att_embed(att))) with zeros on padded regions
 *   p_att16    bf16 [B*L, A]            ctx2att(att_e)
 *   xh16       bf16 [cap+1, B, E+R]     per step: [ x_t | h_{t-1} ]  (operand of the gate GEMM)
 *   s_all      fp32 [cap, B, 5R+A]      per step: i2h(x)+h2h(h) pre-activations | h2att(h)
 *   u_all      fp32 [cap, B, 2R]        per step: a2c(att_res) (+bias)
 *   c_all      fp32 [cap+1, B, R]       cell state, c_all[0] = 0
 *   att_res16  bf16 [cap, B, R]         attention output per step
 *   att_w      fp32 [cap, B, L]         attention weights per step (0 on padded regions)
 *   out16      bf16 [cap, B, R]         dropout(h_t) (operand of the logit GEMM)
 *   z_all      fp32 [cap, B, V1]        vocabulary logits per step
 *   tok_raw / tok_out int64 [cap, B]    sampled id before / after finished-row masking
 *   logp, lse, y_max, y_sum fp32 [cap, B]; unfinished uint8 [cap, B]
 * w_cat16 rows 0..5R-1 = [W_i2h | W_h2h], rows 5R..5R+A-1 = [0 | W_h2att]  (so that one GEMM on
 * [x_t | h_{t-1}] yields the gate pre-activations and att_h).
 *
 * Replaces: AttModel.py:44-51,110-114 (prologue), :465-489 (Attention), :510-531 (Att2in2Core),
 * :74-76 (embed), :87,140,444 (logit + log_softmax), :323-444 (the decode loop of `sample`),
 * :116-141 (the loop of `forward`), gumbel.py:6-30, multinomial.py:4-27.
 */
#define COOPCAP_SAMPLE_GREEDY 0      /* AttModel.py:327-329 */
#define COOPCAP_SAMPLE_MULTINOMIAL 1 /* AttModel.py:332-343: ids only */
#define COOPCAP_SAMPLE_ST_GUMBEL 2   /* gumbel.py:17-30 */
#define COOPCAP_SAMPLE_ST_MULTINOMIAL 3 /* multinomial.py:4-27 */

typedef struct coopcap_speaker {
  /* dimensions */
  int B, L, D, R, E, A, V1;
  int cap;      /* allocated step capacity of the per-step buffers */
  int n_steps;  /* steps to run (<= cap) */
  /* inputs */
  const float* att_feats; /* [B, L, D] */
  const int* att_lens;    /* [B] valid regions per row, or NULL (all L valid) */
  /* parameters */
  const float* embed;     /* [V+2, E] fp32 */
  const float* b_att_embed;
  const float* b_ctx2att;
  const float* b_i2h;
  const float* b_h2h;
  const float* b_h2att;
  const float* b_a2c;
  const float* b_logit;
  const float* w_alpha;   /* [A] */
  const void* w_att_embed16; /* [R, D] */
  const void* w_ctx2att16;   /* [A, R] */
  const void* w_cat16;       /* [5R+A, E+R] */
  const void* w_a2c16;       /* [2R, R] */
  const void* w_logit16;     /* [V1, R] */
  /* randomness: Philox(seed, site stream, element) unless an injected tensor is given */
  uint64_t seed;
  float drop_p;
  const uint8_t* keep_att;   /* [B, L, R] or NULL */
  const uint8_t* keep_embed; /* [cap+1, B, E] or NULL */
  const uint8_t* keep_core;  /* [cap, B, R] or NULL */
  const float* noise;        /* [cap, B, V1] uniforms (gumbel) / Exp(1) draws (multinomial) or NULL */
  /* decode configuration */
  int mode;                  /* COOPCAP_SAMPLE_* */
  float inv_tau;             /* 1/gumbel_temp, 1/multinomial_temp or 1/temperature */
  int64_t start_token;       /* V+1 for `sample` (AttModel.py:324-326), 0 for `forward` (:131) */
  const int64_t* forced;     /* [cap, B] ids that replace the drawn ones (teacher forcing / replay) or NULL */
  /* saved activations / outputs (see layout above) */
  void* att16;
  void* att_e16;
  void* p_att16;
  void* xh16;
  float* s_all;
  float* u_all;
  float* c_all;
  void* att_res16;
  float* att_w;
  void* out16;
  float* z_all;
  int64_t* tok_raw;
  int64_t* tok_out;
  float* logp;
  float* lse;
  float* y_max;
  float* y_sum;
  uint8_t* unfinished;
} coopcap_speaker;

/* att16, att_e16, p_att16 from att_feats (AttModel.py:110-114 / :315-319). */
int coopcap_speaker_prologue_fwd(const coopcap_speaker* ctx, coopcap_stream_t stream);
/* n_steps decode steps: embed -> gates/att_h GEMM -> attention -> a2c GEMM -> LSTM pointwise ->
 * logit GEMM -> sampling (+ next-input gather).  No host synchronisation. */
int coopcap_speaker_decode_fwd(const coopcap_speaker* ctx, coopcap_stream_t stream);

/* d(loss)/d(logits) of the straight-through samplers:
 *   dz = inv_tau * y * (g - <y, g>) on unfinished rows, 0 elsewhere, y = softmax((z+G)*inv_tau)
 * g: [n_steps*B, V1] fp32 (ld ldg) = d(loss)/d(one_hot[:, :V1]); dz16: bf16 [n_steps*B, V1]. */
int coopcap_st_backward(const coopcap_speaker* ctx, const float* g, int64_t ldg, void* dz16,
                        coopcap_stream_t stream);
/* d(loss)/d(logits) of sum_rows coef[row] * log_softmax(z)[row, tok[row]]
 * (REINFORCE: AlternatingJointModel.py:305-309,324; XE: misc/utils.py:49-58 with coef = -mask/sum). */
int coopcap_logp_backward(const coopcap_speaker* ctx, const int64_t* tok, const float* coef,
                          void* dz16, coopcap_stream_t stream);

typedef struct coopcap_speaker_grads {
  /* inputs */
  const void* dz16;     /* bf16 [n_steps*B, V1] */
  /* workspaces */
  float* d_out;         /* [cap*B, R] */
  void* dscat16;        /* bf16 [cap*B, 5R+A] : d(gate pre-acts) | d(att_h) */
  float* d_att_res;     /* [cap*B, R] */
  void* d_att_res16;    /* unused placeholder (must be NULL) */
  float* de;            /* [cap*B, L] d(attention scores) */
  float* dh;            /* [2, B, R] ping-pong */
  float* dc;            /* [2, B, R] ping-pong */
  float* d_x;           /* [cap*B, E] */
  float* d_att_e;       /* [B*L, R] */
  void* d_p_att16;      /* bf16 [B*L, A] */
  void* d_pre16;        /* bf16 [B*L, R] */
  /* outputs: gradients, fp32, same shapes as the reference parameters; written (not accumulated) */
  float* g_embed;       /* [V+2, E]  (must be zero on entry: scatter-add target) */
  float* g_w_att_embed; float* g_b_att_embed;
  float* g_w_ctx2att;   float* g_b_ctx2att;
  float* g_w_cat;       /* [5R+A, E+R] packed like w_cat16 */
  float* g_b_cat;       /* [5R+A]: d b_i2h = d b_h2h = g_b_cat[:5R]; d b_h2att = g_b_cat[5R:] */
  float* g_w_a2c;       float* g_b_a2c;
  float* g_w_logit;     float* g_b_logit;
  float* g_w_alpha;     /* [A] (must be zero on entry) */
} coopcap_speaker_grads;

/* BPTT through the decode loop and the prologue given d(loss)/d(logits). */
int coopcap_speaker_decode_bwd(const coopcap_speaker* ctx, const coopcap_speaker_grads* gr,
                               coopcap_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COOPCAP_H_ */
