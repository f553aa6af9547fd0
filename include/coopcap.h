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

#define COOPCAP_VERSION 200

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
/* Keep [base, base + bytes) resident in L2 for the kernels `stream` launches from now on: reserves
 * min(bytes, device maximum) of the L2 as persisting set-aside (cudaLimitPersistingL2CacheSize, a
 * property of the DEVICE, shared by every stream of the process) and puts a persisting access-policy
 * window over the range on `stream`.  The host side places the bf16 operand copies of the recurrent
 * weights (w_cat16, w_a2c16, w_logit16, w_hh16: 23 MB) in one arena and calls this once: the ~100
 * per-step GEMMs of a training step then find their weights in L2 although every decode step streams
 * ~220 MB (region tensors, logits, gate pre-activations) through it in between.  No reference
 * counterpart (a cache-management call; the reference's cuBLAS path re-reads the weights from DRAM).
 * bytes == 0 removes the window and the set-aside.  granted (may be NULL) receives the bytes
 * actually set aside. */
int coopcap_l2_persist(const void* base, int64_t bytes, int64_t* granted, coopcap_stream_t stream);

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
  int tile_n;  /* 0 = auto, else 64 / 128 / 192 / 256 */
  int backend; /* 0 tcgen05, 1 SIMT cross-check */
} coopcap_gemm_args;

int coopcap_gemm(const coopcap_gemm_args* args, coopcap_stream_t stream);

/* fp32 -> bf16 copy (weights, activations); optional transposed copy dst_t [cols, rows]. */
int coopcap_cast_bf16(const float* src, int64_t rows, int64_t cols, int64_t ld_src, void* dst,
                      int64_t ld_dst, void* dst_t, int64_t ld_dst_t, coopcap_stream_t stream);


/* ---- speaker (Att2in2) -----------------------------------------------------------------------
 * One context struct describes a whole speaker pass over B rows: dimensions, inputs, packed
 * parameters, RNG configuration, and every activation the backward pass needs.  All buffers are
 * caller-allocated; the library only launches kernels (no allocation, no synchronisation).
 *
 * Region tensors are PACKED: only valid regions are stored, row b owning packed rows
 * att_off[b] .. att_off[b+1]-1 (att_off == NULL: every row has L regions).  This is the varlen
 * restatement of pack_wrapper (AttModel.py:31-51) + the mask-renormalised softmax (:480-483).
 *
 * Layout in HBM (row-major, sizes in elements; NL = number of valid regions in the batch):
 *   att16      bf16 [NL, D]             cast + pack of att_feats (operand of att_embed fwd / wgrad)
 *   att_e16    bf16 [NL, R]             dropout(relu(att_embed(att)))
 *   p_att16    bf16 [NL, A]             ctx2att(att_e)
 *   xh16       bf16 [cap+1, B, E+R]     per step: [ x_t | h_{t-1} ]  (operand of the gate GEMM)
 *   s_all      fp32 [cap, B, 5R+A]      per step: i2h(x)+h2h(h) pre-activations | h2att(h)
 *   u_all      fp32 [cap, B, 2R]        per step: a2c(att_res) (+bias)
 *   c_all      fp32 [cap+1, B, R]       cell state, c_all[0] = 0
 *   att_res16  bf16 [cap, B, R]         attention output per step
 *   att_res32  fp32 [cap, B, R]         the same in fp32 (optional; single-pass attention backward)
 *   att_w      fp32 [cap, NL]           attention weights per step (packed like the regions)
 *   out16      bf16 [cap, B, R]         dropout(h_t) (operand of the logit GEMM)
 *   z16_all    fp16 [cap, B, V1]        vocabulary logits per step.  The sampler runs on the fp32
 *                                       accumulators inside the logit GEMM's epilogue (logit_sample.cuh);
 *                                       this half-precision copy is what backward rebuilds y from
 *   ls_part    fp32 [B, 4*ceil(V1/256), 8]  per-(row, 64-column group) partials of one step's epilogue
 *   z_tgt      fp32 [B]                 raw logit of the forced id of one step
 *   tok_raw    int64 [cap, B]           id drawn from z_t (before forcing / finished-row masking)
 *   tok_out    int64 [cap, B]           id after forcing and finished-row masking (AttModel.py:409)
 *   tok_fed    int64 [cap+1, B]         id embedded as the input of step t (tok_fed[0] = start)
 *   logp, lse, y_max, y_sum fp32 [cap, B]; unfinished uint8 [cap, B]
 * w_cat16 rows 0..5R-1 = [W_i2h | W_h2h], rows 5R..5R+A-1 = [0 | W_h2att]  (one GEMM on
 * [x_t | h_{t-1}] yields the gate pre-activations and att_h);  b_cat = [b_i2h+b_h2h | b_h2att].
 *
 * Replaces: AttModel.py:44-51,110-114 (prologue), :465-489 (Attention), :510-531 (Att2in2Core),
 * :74-76 (embed), :87,140,444 (logit + log_softmax), :323-444 (the decode loop of `sample`),
 * :116-141 (the loop of `forward`), gumbel.py:6-30, multinomial.py:4-27.
 */
#define COOPCAP_SAMPLE_GREEDY 0         /* AttModel.py:327-329 */
#define COOPCAP_SAMPLE_MULTINOMIAL 1    /* AttModel.py:332-343: ids only */
#define COOPCAP_SAMPLE_ST_GUMBEL 2      /* gumbel.py:17-30 */
#define COOPCAP_SAMPLE_ST_MULTINOMIAL 3 /* multinomial.py:4-27 */
#define COOPCAP_SAMPLE_NONE 4           /* teacher forcing: only lse / logp of the forced id */
#define COOPCAP_SAMPLE_PS_GUMBEL 5      /* gumbel_softmax.py:17-42 (partial sampling) */
#define COOPCAP_SAMPLE_PS_MULTINOMIAL 6 /* multinomial_soft.py:5-35 (partial sampling) */

/* fp32 master parameters -> packed bf16 operand copies (run after every optimizer step). */
typedef struct coopcap_speaker_pack {
  int D;
  int R;
  int E;
  int A;
  int V1;
  const float* w_att_embed; /* [R, D]   att_embed.0.weight */
  const float* w_ctx2att;   /* [A, R]   ctx2att.weight */
  const float* w_i2h;       /* [5R, E]  core.i2h.weight */
  const float* w_h2h;       /* [5R, R]  core.h2h.weight */
  const float* w_h2att;     /* [A, R]   core.attention.h2att.weight */
  const float* w_a2c;       /* [2R, R]  core.a2c.weight */
  const float* w_logit;     /* [V1, R]  logit.weight */
  const float* b_i2h;
  const float* b_h2h;
  const float* b_h2att;
  void* w_att_embed16;
  void* w_ctx2att16;
  void* w_cat16;            /* [5R+A, E+R] */
  void* w_a2c16;
  void* w_logit16;
  float* b_cat;             /* [5R+A] */
} coopcap_speaker_pack;

int coopcap_speaker_pack_weights(const coopcap_speaker_pack* p, coopcap_stream_t stream);
/* Weight re-pack launches: 1 (default) = the casts of one coopcap_*_pack_weights call run as one
 * fused launch, 0 = one launch per tensor.  Process-wide host-side switch (no device state); the
 * zero-copy upload path (coopcap_pack_att_from_host) turns it off because its reader kernel shares
 * the SMs with the re-pack.  No reference counterpart. */
int coopcap_set_cast_multi(int on);

typedef struct coopcap_speaker {
  /* dimensions */
  int B;
  int L;        /* padded region width of att_feats */
  int D;
  int R;
  int E;
  int A;
  int V1;
  int NL;       /* total valid regions (== B*L when att_off is NULL) */
  int cap;      /* allocated step capacity of the per-step buffers */
  int n_steps;  /* steps to run (<= cap) */
  /* inputs */
  const float* att_feats; /* [B, L, D] */
  const int* att_off;     /* [B+1] packed-row offsets, or NULL */
  const int* att_order;   /* [B] row ids by decreasing region count, or NULL (identity): only a
                             schedule hint, the attention kernels deal rows to SMs in this order */
  int att_prepacked;      /* 1: att16 already holds the packed bf16 regions (att_feats unused) */
  /* parameters */
  const float* embed;     /* [V+2, E] fp32 */
  const float* b_att_embed;
  const float* b_ctx2att;
  const float* b_cat;
  const float* b_a2c;
  const float* b_logit;
  const float* w_alpha;   /* [A] alpha_net.weight (its bias cancels in the softmax) */
  const void* w_att_embed16;
  const void* w_ctx2att16;
  const void* w_cat16;
  const void* w_a2c16;
  const void* w_logit16;
  /* randomness: Philox(seed, site stream, element) unless an injected tensor is given */
  uint64_t seed;
  float drop_p;
  const uint8_t* keep_att;   /* [NL, R] packed, or NULL */
  const uint8_t* keep_embed; /* [>= n_steps, B, E] or NULL */
  const uint8_t* keep_core;  /* [cap, B, R] or NULL */
  const float* noise;        /* [cap, B, V1] uniforms (gumbel) / Exp(1) draws (multinomial) or NULL */
  /* decode configuration */
  int mode;                  /* COOPCAP_SAMPLE_* */
  float inv_tau;             /* 1/gumbel_temp, 1/multinomial_temp or 1/temperature */
  int64_t start_token;       /* V+1 for `sample` (AttModel.py:324-326), 0 for `forward` (:131) */
  const int64_t* start_tokens; /* [B] per-row start ids (seq[:,0], AttModel.py:131) or NULL -> start_token */
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
  void* z16_all;
  float* ls_part;
  float* z_tgt;
  int64_t* tok_raw;
  int64_t* tok_out;
  int64_t* tok_fed;
  float* logp;
  float* lse;
  float* y_max;
  float* y_sum;
  uint8_t* unfinished;
  /* caption summary written by coopcap_speaker_decode_fwd after the last step:
   * n_out[0] = n = output width of `sample` (AttModel.py:407-408), cap_len[b] = number of valid
   * positions of [BOS, w_1..w_n] under _masks = [1,1,(w_1>0),..] (AlternatingJointModel.py:353-355) */
  int* n_out;                /* [1] */
  int* cap_len;              /* [B] */
  /* partial-sampling modes (COOPCAP_SAMPLE_PS_*; AttModel.py:367-399,425-434): step t emits the
   * vector v_t = one_hot(id) on rows with part_u < ps_prob and the relaxed sample y on the others
   * (y = softmax((z+G)/tau) resp. exp(log_softmax(z)/tau)); the next input is
   * dropout(relu(v_t . embed)) instead of an embedding lookup.  After that GEMM, finished rows of
   * v_t are replaced by the EOS one-hot (:428-432), which is what `sample` returns. */
  float ps_prob;             /* prob_gumbel_softmax / prob_multinomial_soft; <= 0: every row stays soft */
  const float* part_u;       /* [cap, B] injected uniforms or NULL -> Philox */
  void* soft16;              /* bf16 [cap, B, V1]  v_t (the BOS column V+1 is always 0 and not stored) */
  uint8_t* ps_sel;           /* [cap, B] 1 = row emitted the hard one-hot */
  const void* w_embed16;     /* bf16 [>= V1, E] copy of `embed` */
  /* scheduled sampling of AttModel.forward (:119-131): with probability ss_prob per row the next
   * input is the id drawn from softmax(z_t) (mode COOPCAP_SAMPLE_MULTINOMIAL) instead of
   * forced[t]; logp stays the log-probability of forced[t] (the XE target). */
  float ss_prob;
  const float* ss_u;         /* [cap, B] injected uniforms or NULL -> Philox */
  /* decoding_constraint (AttModel.py:437-442): the logit of the previously emitted id is -inf */
  int no_repeat;
  /* COOPCAP_SAMPLE_ST_GUMBEL only: z16_all receives the PERTURBED logits z + G instead of z.  The
   * relaxed sample y = softmax((z+G)/tau) -- all the straight-through backward needs (gumbel.py:13-30,
   * SURVEY.md A.3) -- is then rebuilt from one fp16 read, with no noise regenerated (st_bwd: 95 -> ~50 us
   * per 4096 rows).  logp / lse are still exact outputs of the forward pass, but softmax(z) can no
   * longer be rebuilt, so coopcap_logp_backward refuses such a context: callers that differentiate
   * the sampled ids' log-probabilities (the CIDEr term) leave this 0. */
  int store_perturbed;
  /* fp32 copy of the attention output, [cap, B, R], or NULL.  When present (and A == R == 512) the
   * per-step attention backward is a single pass over (att_e, p_att): the softmax backward's mean
   * sum_l w_l <d_att_res, att_e_l> equals <d_att_res, att_res> (csrc/attention.cuh, v5). */
  float* att_res32;
} coopcap_speaker;

/* att16, att_e16, p_att16 from att_feats (AttModel.py:110-114 / :315-319). */
int coopcap_speaker_prologue_fwd(const coopcap_speaker* ctx, coopcap_stream_t stream);
/* n_steps decode steps: embed -> gates/att_h GEMM -> attention -> a2c GEMM -> LSTM pointwise ->
 * logit GEMM with the sampler in its epilogue -> per-row finish (+ next-input gather).  No host
 * synchronisation. */
int coopcap_speaker_decode_fwd(const coopcap_speaker* ctx, coopcap_stream_t stream);

/* d(loss)/d(logits) of the straight-through samplers for all n_steps:
 *   g  = demb[t+1] . W_emb^T                  (listener-embedding dgrad, VSEFCModel.py:104)
 *   dz = inv_tau * y * (g - <y, g>) on unfinished rows, 0 elsewhere, y = softmax((z+G)*inv_tau)
 * demb16: bf16 [n_steps*B, E] gradient w.r.t. the listener's word embedding at caption positions
 * 1..n_steps; w_emb16: bf16 [>=V1, E]; g_ws: bf16 [g_chunk_steps*B, V1] workspace (g is formed for
 * g_chunk_steps steps per launch: larger GEMM tiles, scratch mostly L2-resident);
 * dz16: bf16 [n_steps*B, V1]. */
int coopcap_st_backward(const coopcap_speaker* ctx, const void* demb16, const void* w_emb16,
                        void* g_ws, int g_chunk_steps, void* dz16, coopcap_stream_t stream);
/* Same with a dense upstream gradient g = d(loss)/d(one_hots[:, :, :V1]) given explicitly
 * (fp32 [n_steps*B, ldg]); used when foreign code consumed the dense one-hot tensor. */
int coopcap_st_backward_dense(const coopcap_speaker* ctx, const float* g, int64_t ldg, void* dz16,
                              coopcap_stream_t stream);
/* d(loss)/d(logits) of sum_{t,b} coef[t,b] * log_softmax(z[t,b])[tok[t,b]]
 * (REINFORCE: AlternatingJointModel.py:305-309,324; XE: misc/utils.py:49-58 with coef = -mask/sum).
 * tok: int64 [n_steps, B]; coef: fp32 [n_steps, B]. */
int coopcap_logp_backward(const coopcap_speaker* ctx, const int64_t* tok, const float* coef,
                          void* dz16, coopcap_stream_t stream);
/* Same, ADDED to the gradient already in dz16: a second loss term on the log-probabilities of a
 * pass whose logits already carry another gradient (the CIDEr term of traditional_cider next to
 * the straight-through listener gradient, AlternatingJointModel.py:409-431,490-503). */
int coopcap_logp_backward_acc(const coopcap_speaker* ctx, const int64_t* tok, const float* coef,
                              void* dz16, coopcap_stream_t stream);

typedef struct coopcap_speaker_grads {
  /* input */
  const void* dz16;     /* bf16 [n_steps*B, V1] (partial-sampling passes: a workspace, written) */
  /* workspaces */
  float* d_out;         /* [max(cap,2)*B, R]  d(loss)/d(dropout(h_t)) from the logit layer (reused as scratch) */
  void* dscat16;        /* bf16 [cap*B, 5R+A] : d(gate pre-acts) | d(att_h) */
  float* d_att_res;     /* [cap*B, R] */
  float* de;            /* [cap, NL] d(attention scores) */
  float* d_xh;          /* [cap*B, E+R] : d(x_t) | d(h_{t-1}) */
  float* dc;            /* [2, B, R] ping-pong */
  float* d_att_e;       /* [NL, R] */
  void* d_p_att16;      /* bf16 [NL, A] */
  void* d_pre16;        /* bf16 [NL, R] */
  /* outputs: gradients, fp32, reference parameter shapes; written (overwritten), except g_embed
   * which is accumulated into (the caller zeroes it) */
  float* g_embed;       /* [V+2, E] */
  float* g_w_att_embed;
  float* g_b_att_embed;
  float* g_w_ctx2att;
  float* g_b_ctx2att;
  float* g_w_i2h;
  float* g_w_h2h;
  float* g_b_gates;     /* [5R]: d b_i2h = d b_h2h */
  float* g_w_h2att;
  float* g_b_h2att;
  float* g_w_a2c;
  float* g_b_a2c;
  float* g_w_logit;
  float* g_b_logit;
  float* g_w_alpha;     /* [A] */
  /* partial-sampling modes only (COOPCAP_SAMPLE_PS_*): the gradient of v_t has two sources, the
   * consumer of the emitted vectors (listener) and the next input x_{t+1} = relu(v_t . embed), so
   * d(loss)/d(logits) is formed step by step inside the BPTT loop and dz16 is a WORKSPACE
   * (bf16 [n_steps*B, V1]) instead of an input.  Upstream gradient, one of:
   *   factored: ps_demb16 bf16 [n_steps*B, E] (d loss / d listener word embedding at caption
   *             positions 1..n_steps) and ps_w_emb16 bf16 [>= V1, E]; ps_g fp32 [B, V1] scratch
   *   dense   : ps_demb16 NULL; ps_g fp32 [n_steps*B, ps_ldg] = d loss / d v, modified in place */
  const void* ps_demb16;
  const void* ps_w_emb16;
  float* ps_g;
  int64_t ps_ldg;
  void* ps_dpre16;      /* bf16 [n_steps*B, E] workspace: d(loss)/d(v_{t-1} . embed) */
  /* 0 = the whole backward pass; 1 = only the vocabulary layer (d_out, g_w_logit, g_b_logit);
   * 2 = everything after it.  A data-parallel caller runs 1, starts the exchange of the 19.5 MB
   * logit gradients between the ranks, then runs 2 (the BPTT loop hides the exchange).  The
   * partial-sampling modes form the logit gradients inside the loop: they take 0 only. */
  int phase;
} coopcap_speaker_grads;

/* BPTT through the decode loop and the prologue given d(loss)/d(logits). */
int coopcap_speaker_decode_bwd(const coopcap_speaker* ctx, const coopcap_speaker_grads* gr,
                               coopcap_stream_t stream);

/* ---- listener (VSEFC) ------------------------------------------------------------------------
 * Image encoder, GRU caption encoder over index captions, cosine score matrix and max-violation
 * hinge loss, forward and backward.  Captions are time-major ids tok [S, B] with per-row valid
 * lengths len [B] (= sum(mask > 0), VSEFCModel.py:84); the packed GRU (:108-112) is restated as a
 * masked update so no sort / unsort is needed.
 *
 * Layout in HBM:
 *   fc16     bf16 [B, F]           cast of fc_feats
 *   img_pre  fp32 [B, M]           fc W^T + b              (VSEFCModel.py:44)
 *   im       fp32 [B, M]           l2norm(img_pre)         (:12-17,47-48)
 *   emb16    bf16 [S, B, E]        Embedding[tok]          (:102-106)
 *   gi_all   fp32 [S, B, 3M]       x W_ih^T + b_ih for all positions
 *   gh       fp32 [B, 3M]          h W_hh^T + b_hh (per-step scratch)
 *   gates    fp32 [S, B, 4M]       r | z | n | (W_hn h + b_hn) per step (saved for backward)
 *   h32      fp32 [S+1, B, M]      hidden state, h32[0] = 0
 *   h16      bf16 [S+1, B, M]
 *   cap_pre  == h32[S]             state after len steps (pool 'last', :128)
 *   cap      fp32 [B, M]           l2norm(cap_pre)
 *   scores   fp32 [B, B]           im . cap^T              (:143-146)
 *   cost_s, cost_im fp32 [B]; arg_s, arg_im int32 [B]   max-violation terms (:176-193)
 *   loss_rows fp32 [B] (whole_batch vector, :197-207); loss fp32 [1] (their sum)
 */
typedef struct coopcap_listener_pack {
  int F;
  int M;
  int E;
  int V2;
  const float* w_img;   /* [M, F]  img_enc.fc.weight */
  const float* w_ih;    /* [3M, E] txt_enc.rnn.weight_ih_l0 */
  const float* w_hh;    /* [3M, M] txt_enc.rnn.weight_hh_l0 */
  const float* w_emb;   /* [V2, E] txt_enc.embed.weight */
  void* w_img16;
  void* w_ih16;
  void* w_hh16;
  void* w_emb16;
} coopcap_listener_pack;

int coopcap_listener_pack_weights(const coopcap_listener_pack* p, coopcap_stream_t stream);

typedef struct coopcap_listener {
  int B;
  int S;       /* caption positions (incl. BOS) */
  int F;
  int M;
  int E;
  int V2;
  float margin;
  int only_one_retrieval; /* 0 off, 1 image, 2 caption (VSEFCModel.py:202-207) */
  int no_imgnorm;
  /* inputs */
  const float* fc_feats;  /* [B, F] */
  const int64_t* tok;     /* [S, B] time-major ids */
  const int* len;         /* [B] valid positions per row */
  /* parameters */
  const float* w_emb;     /* [V2, E] fp32 */
  const float* b_img;
  const float* b_ih;
  const float* b_hh;
  const void* w_img16;
  const void* w_ih16;
  const void* w_hh16;
  /* saved activations / outputs */
  void* fc16;
  float* img_pre;
  float* im;
  void* emb16;
  float* gi_all;
  float* gh;
  float* gates;
  float* h32;
  void* h16;
  float* cap;
  float* scores;
  float* cost_s;
  float* cost_im;
  int* arg_s;
  int* arg_im;
  float* loss_rows;
  float* loss;
  /* variants (0 / NULL = the run_joint.sh defaults) */
  int emb_given;      /* 1: emb16 was filled by the caller (dense one-hot / soft caption vectors,
                         VSEFCModel.py:102-104); tok is not read */
  int pool_type;      /* 0 last (:128), 1 masked mean (:115-119), 2 masked max (:120-126) */
  int use_abs;        /* vse_use_abs: |l2norm(.)| on both encoders (:50-52,137-139) */
  int sum_violation;  /* 1: vse_max_violation = 0, mean over the negatives (:190-193) */
  float* cap_pre;     /* [B, M] pooled state before l2norm (pool mean / max; 'last' uses h32[S]) */
  int* pool_arg;      /* [B, M] arg-max step of the max pool */
} coopcap_listener;

int coopcap_listener_fwd(const coopcap_listener* ctx, coopcap_stream_t stream);

/* Caption embedding of dense caption vectors (VSEFCModel.py:102-104, `seqs.dim() > 2`) for the
 * partial-sampling joint step: emb16[0] = bf16(w_emb[bos_id]) on every row (the BOS one-hot that
 * AlternatingJointModel.py:356-370 prepends), emb16[1 + t] = v_t . w_emb[:V1]  for t < n.
 * soft16: bf16 [n, B, V1]; w_emb fp32 [V2, E]; w_emb16 bf16 [V2, E]; emb16: bf16 [n+1, B, E]. */
int coopcap_caption_embed_dense(const void* soft16, const float* w_emb, const void* w_emb16, int n,
                                int B, int V1, int E, int64_t bos_id, void* emb16,
                                coopcap_stream_t stream);
/* Its weight gradient: g_w_emb[:V1] = sum_t v_t^T . demb[1 + t], g_w_emb[bos_id] = sum_b demb[0, b]
 * (positions beyond a row's length carry zero demb).  g_w_emb: fp32 [V2, E], other rows untouched. */
int coopcap_caption_embed_dense_bwd(const void* soft16, const void* demb16, int n, int B, int V1,
                                    int E, int64_t bos_id, float* g_w_emb, coopcap_stream_t stream);

typedef struct coopcap_listener_grads {
  /* upstream gradient: d(total)/d(loss) scalar (device, [1]) or per-row vector [B]; exactly one */
  const float* g_loss;
  const float* g_rows;
  int need_param_grads;  /* 0: only demb16 (speaker turn with a frozen listener) */
  /* workspaces */
  float* d_im;          /* [B, M] */
  float* d_cap;         /* [B, M] */
  float* dh;            /* [B, M] running d(h) */
  void* d_img_pre16;    /* bf16 [B, M] */
  void* d_gi16;         /* bf16 [S, B, 3M] */
  void* d_gh16;         /* bf16 [S, B, 3M] */
  /* outputs */
  void* demb16;         /* bf16 [S, B, E]  d(loss)/d(word embedding input) */
  float* g_w_img;
  float* g_b_img;
  float* g_w_emb;       /* [V2, E]  accumulated into (caller zeroes) */
  float* g_w_ih;
  float* g_w_hh;
  float* g_b_ih;
  float* g_b_hh;
  /* variants */
  float* d_pool;        /* [B, M] workspace: d(loss)/d(pooled state), pool mean / max only */
  float* d_scores;      /* [B, B] workspace: d(loss)/d(scores), sum-violation hinge only */
} coopcap_listener_grads;

int coopcap_listener_bwd(const coopcap_listener* ctx, const coopcap_listener_grads* gr,
                         coopcap_stream_t stream);

/* ---- optimizer (optimizer.py:233-242 + misc/utils.py:65-69 + torch.optim.Adam) ---------------
 * One pass over a flat fp32 bucket: g *= grad_scale (1/world_size after the all-reduce);
 * g = clamp(g, -clip, clip); Adam(lr, beta1, beta2, eps, weight_decay) with bias correction for
 * `step` (1-based).  clip <= 0 disables clamping.  The hyper-parameters are doubles, as in
 * torch.optim.Adam, whose scalar factors (1 - beta, lr / (1 - beta1^step), sqrt(1 - beta2^step)) are
 * formed in double on the host and only then rounded to fp32: the update is bit-comparable. */
int coopcap_clamp_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       int64_t n, double grad_scale, double clip, double lr, double beta1,
                       double beta2, double eps, double weight_decay, int step,
                       coopcap_stream_t stream);

/* ---- host -> device staging (train.py:162-178 `load_data` / misc/utils.py:72-87 `var_wrapper`) ----
 * The loader zero-pads att_feats to the longest image of the batch (dataloader.py:220-229); only
 * the valid prefix of every row needs to cross PCIe.  Copies lens_host[b] * unit_bytes bytes of row
 * b (row pitch row_stride_bytes in both buffers) from pinned host memory, one cudaMemcpyAsync per
 * row on `stream`.  lens_host is a HOST array.  Padded tails of dst are left untouched (the kernels
 * never read them: they work on the packed valid regions). */
int coopcap_h2d_ragged_rows(void* dst, const void* src_host, const int* lens_host, int B,
                            int64_t row_stride_bytes, int64_t unit_bytes, coopcap_stream_t stream);

/* Zero-copy variant: a small persistent grid reads the valid regions of a PINNED host tensor
 * att_feats [B, L, D] fp32 straight over PCIe and writes the packed bf16 operand att16 [NL, D]
 * (the same result coopcap_speaker_prologue_fwd's pack step produces from a device tensor), so the
 * padded fp32 tensor never exists in HBM.  att_off is a DEVICE array [B+1] (or NULL: all L valid).
 * Meant to run on a copy stream while the previous step computes; `ctas` bounds the SMs it uses
 * (0 = default 64; 128-thread CTAs that co-reside with the GEMM CTAs).  Pass the result as coopcap_speaker.att16 with att_prepacked = 1. */
int coopcap_pack_att_from_host(const float* att_feats_pinned, const int* att_off, int B, int L, int D,
                               int NL, void* att16, int ctas, coopcap_stream_t stream);

/* Host-side variant of the batch staging (reference: `load_data` / `utils.var_wrapper(...).cuda()`,
 * train.py:162-178, misc/utils.py:72-87): worker threads of the library pack the valid regions of a HOST tensor
 * att_feats [B, L, D] fp32 into a (pinned) HOST staging buffer att16_host [NL, D] bf16 -- same
 * round-to-nearest-even as the device pack, so the operand is bit-identical -- which the caller
 * then moves with one DMA copy: a quarter of the padded fp32 bytes cross PCIe and no SM is used.
 * att_off_host is a HOST array [B+1] (or NULL: all L valid); nthreads <= 0 picks hardware threads - 1.
 * coopcap_host_pack_start returns a job id >= 0 immediately (or a negative error code); the buffers
 * must stay valid until coopcap_host_pack_wait(job) has returned. */
int coopcap_host_pack_start(const float* att_feats, const int* att_off_host, int B, int L, int D,
                            void* att16_host, int nthreads);
int coopcap_host_pack_wait(int job);

/* ---- device-resident feature store (dataloader.py:137-160,220-229 + train.py:162-178) ------------
 * The reference's loader reads an image's region features by image index, zero-pads the batch and
 * `load_data` copies it to the GPU every step.  With the whole feature set resident in HBM as
 * packed bf16 rows (store_att16 [total_regions, D], image i owns rows store_off[i]..store_off[i+1]),
 * a step ships only indices: this call gathers the images ix[0..B) into the packed operand
 * att16_out [NL, D] (row b's regions at att_off[b]..att_off[b+1], the layout of
 * coopcap_speaker.att16 with att_prepacked = 1) and, when fc_out != NULL, their fc vectors
 * store_fc [n_img, F] fp32 into fc_out [B, F].  ix, store_off, att_off are DEVICE arrays; att_off
 * must be the running sum of the gathered images' region counts (the host knows them). */
int coopcap_store_gather(const void* store_att16, const int64_t* store_off, const float* store_fc,
                         int64_t n_img, const int64_t* ix, int B, int D, int F, const int* att_off,
                         void* att16_out, float* fc_out, coopcap_stream_t stream);

/* ---- beam search (AttModel.py:150-289 `sample_beam`, evaluation) ----------------------------------
 * The reference decodes one image at a time in Python: beam_size rows through the core, a CPU sort
 * of the [beam, V+1] log-probabilities, a Python list of candidates, per-beam state copies.  Here
 * all images advance together: rows are BEAM-MAJOR (row = slot * n_img + image), so every per-step
 * kernel of the decode loop is reused as is (the attention kernel runs once per slot over the
 * n_img images with the context's own region offsets), and one CTA per image does the merge.
 * Semantics kept exactly, quirks included (oracle/speaker.py `sample_beam` lists them): the first
 * merge looks at slot 0 only; candidates are the top beam_size words of every slot, listed
 * word-rank-major, stable-sorted by total log-probability; a slot that emitted the end token is
 * recorded but NOT retired; at step T every slot is recorded; the recorded score is a VIEW of the
 * running sum, so the done list is ranked by each slot's FINAL sum (ties: recording order).
 * `ctx` is an ordinary speaker context for the n_img images (ctx->B = n_img) whose prologue has run
 * (coopcap_speaker_prologue_fwd); only its parameters, att_e16 / p_att16 / att_off / att_order are
 * read.  Evaluation mode: no dropout. */
typedef struct coopcap_beam {
  int beam_size;          /* 1 .. 8 */
  int no_repeat;          /* decoding_constraint (:204-207): -inf on the slot's previous word, t > 1 */
  int T;                  /* seq_length */
  int reserved;
  /* step buffers, rows = beam_size * n_img */
  void* xh16;             /* bf16 [2, rows, E+R] */
  float* c2;              /* [2, rows, R] */
  float* s_t;             /* [rows, 5R+A] */
  float* u_t;             /* [rows, 2R] */
  void* att_res16;        /* bf16 [rows, R] */
  float* att_w;           /* [beam_size, NL] */
  void* h_stage16;        /* bf16 [rows, R] */
  float* c_stage;         /* [rows, R] */
  float* logits;          /* [rows, V1] */
  int* parent;            /* [rows] */
  int64_t* tok;           /* [rows] */
  int64_t* hist_seq;      /* [2, n_img, T, beam_size], zero-initialised by the call */
  float* hist_lp;         /* [2, n_img, T, beam_size] */
  float* beam_sum;        /* [n_img, beam_size] */
  /* replay aid for parity tests: decisions to APPLY at merge step t = 1..T (NULL: the kernel's own)
   * and the decisions the kernel WOULD have taken, both [T, n_img, beam_size] */
  const int* forced_parent;
  const int64_t* forced_tok;
  int* raw_parent;
  int64_t* raw_tok;
  /* outputs */
  int64_t* done_seq;      /* [n_img, beam_size*T, T] every recorded beam, in recording order */
  float* done_lp;         /* [n_img, beam_size*T, T] */
  int* done_slot;         /* [n_img, beam_size*T] */
  float* done_p_rec;      /* [n_img, beam_size*T] sum at recording time */
  float* done_p;          /* [n_img, beam_size*T] the score the reference ranks by (slot's final sum) */
  int* done_n;            /* [n_img] */
  int64_t* seq;           /* [n_img, T] first of the ranked list (:286) */
  float* seq_logp;        /* [n_img, T] (:287) */
} coopcap_beam;
int coopcap_speaker_beam_fwd(const coopcap_speaker* ctx, const coopcap_beam* beam, coopcap_stream_t stream);

/* ---- retrieval evaluation (eval_utils.py:545-595 `i2t`, :598-720 `t2i`) ------------------------------
 * Rank of the correct item for every query of a [n_query, n_cand] fp32 score matrix: the number
 * of candidates scoring strictly higher than the best-scoring correct one (numpy argsort
 * descending + np.where, ties aside).  Query q's correct candidates are
 * first[q] .. first[q] + count - 1 (i2t: the image's 5 captions; t2i: the caption's image).
 * ranks / top1: int32 [n_query]. */
int coopcap_retrieval_ranks(const float* scores, int64_t ld, int n_query, int n_cand, const int* first,
                            int count, int* ranks, int* top1, coopcap_stream_t stream);
/* scores[q, n] = <queries[q, :], cands[n, :]> in fp32 FMAs (np.dot of float32 arrays,
 * eval_utils.py:573,631): the ranks above must not depend on bf16 / tf32 rounding.
 * queries [n_query, K], cands [n_cand, K], scores [n_query, ld]. */
int coopcap_retrieval_scores(const float* queries, const float* cands, int n_query, int n_cand, int K,
                             float* scores, int64_t ld, coopcap_stream_t stream);

/* ---- CIDEr-D self-critical reward (misc/rewards.py:34-71, ciderD_scorer.py:13-28,105-215) --------
 * The reference scores the sampled and the greedy captions of a batch against each image's
 * ground-truth captions on the host (Python dictionaries of word n-grams, float64) every step.
 * Here the ids stay on the device.  An n-gram (n = 1..4) of ids < 65535 is one exact 64-bit key
 * (16 bits per id + 1, 0 = absent), so nothing is hashed lossily:
 *   cook    one warp per caption: words up to AND INCLUDING the first 0 (rewards.py:26-32), unique
 *           n-grams with counts in first-occurrence order (precook's dictionary order)
 *   df      "corpus" mode (opts.py:27 default): number of ENTRIES (hypotheses) whose image's
 *           references contain the n-gram (ciderD_scorer.py:105-118), built in an open-addressing
 *           table of exact keys; cached mode (--cached_tokens file): the caller uploads the table
 *   vec     tf-idf weight cnt * (log_ref_len - log(max(1, df))) and per-order norms, float64
 *   score   clipped dot products, cosine normalisation, Gaussian length penalty on the bigram
 *           counts (sic), mean over orders / references, x 10  (ciderD_scorer.py:147-203)
 *   finish  reward = score(sampled) - score(greedy) (or the sampled score alone) as fp32, and the
 *           per-token REINFORCE coefficients of traditional_cider
 *           (AlternatingJointModel.py:409-431): coef[t,b] = -reward[b] * mask[b,t] / sum(mask),
 *           mask[b,t] = 1 for t <= (leading non-zero ids of row b) and t < n, n = caption width.
 * Sums run in the reference's order, so scores agree with it to float64 rounding of log/pow.
 * Captions c = 0 .. n_sets*B-1 are the hypotheses (set-major), then the n_ref references. */
typedef struct coopcap_cider {
  int B;                 /* rows per hypothesis set */
  int n_sets;            /* 1 (sampled only) or 2 (sampled, greedy) */
  int T;                 /* steps in the hypothesis arrays, <= 16 */
  int W;                 /* width of a reference caption, <= 16 */
  int n_img;             /* images in the batch */
  int n_ref;             /* reference captions of all images */
  int corpus;            /* 1: document frequencies from this batch; 0: df table given */
  int differenced;       /* 1: reward = sampled - greedy (use_gen_cider_scores == 0); 0: sampled */
  double log_ref_len;    /* cached mode: log(ref_len of the table); corpus mode: ignored */
  const int64_t* hyp0;   /* int64 [T, B] time-major ids of the sampled captions */
  const int64_t* hyp1;   /* int64 [T, B] greedy captions (n_sets == 2) or NULL */
  const int64_t* refs;   /* int64 [n_ref, W] ground-truth captions, 0-padded */
  const int* ref_off;    /* [n_img + 1]: image i owns refs ref_off[i] .. ref_off[i+1]-1 */
  const int* row_img;    /* [B]: image of row b (b / seq_per_img in the reference) */
  uint64_t* df_keys;     /* [df_cap] open-addressing table, 0 = empty (corpus mode: workspace) */
  float* df_val;         /* [df_cap] document frequencies (integers, exact in fp32) */
  int df_cap;            /* power of two; corpus mode needs >= 2 * 58 * n_ref */
  int reserved;
  /* workspaces, C = n_sets*B + n_ref captions */
  uint64_t* ng_key;      /* [C, 64] unique n-gram keys, order n at slots 16n .. 16n+15 */
  int* ng_cnt;           /* [C, 64] term frequencies */
  int* ng_n;             /* [C, 4] unique n-grams per order */
  int* ng_len;           /* [C] "length" = number of bigrams (ciderD_scorer.py:143-144) */
  double* ng_w;          /* [C, 64] tf-idf weights */
  double* ng_norm;       /* [C, 4] */
  int* img_rows;         /* [n_img] rows per image */
  /* outputs */
  double* scores;        /* [n_sets * B] CIDEr-D of every hypothesis */
  float* reward;         /* [B] */
  float* coef;           /* [T, B] or NULL */
  double* stats;         /* [4]: mean reward, mean greedy score, sum(mask), mean sampled score */
} coopcap_cider;
int coopcap_cider_reward(const coopcap_cider* ctx, coopcap_stream_t stream);
/* Slot of `key` in a table of `cap` (power of two) entries before probing; the caller that
 * builds a cached table on the host probes linearly from here (same function as the device). */
uint64_t coopcap_cider_hash(uint64_t key);

/* ---- instrumentation ---------------------------------------------------------------------------
 * coopcap_launch_count: kernels launched by this library since load (all streams).
 * coopcap_prof_enable(1, stream): start an event timeline on `stream` (one event after every
 * launch; the library's launches on one stream run back to back, so the gap between consecutive
 * events is that launch's duration); coopcap_prof_report sums per kernel class (index = ProfKind in
 * csrc/common.cuh: 0 misc, 1 gemm, 2 att_fwd, 3 att_bwd, 4 att_deferred, 5 lstm, 6 sample,
 * 7 st_bwd, 8 logp_bwd, 9 gru, 10 hinge, 11 reduce, 12 pack, 13 adam, 14 logit_sample = the logit
 * GEMM with the sampler in its epilogue) the elapsed ms, the
 * algorithmic FLOPs / bytes the launches declared, and the launch count, then clears the timeline. */
long long coopcap_launch_count(void);
int coopcap_prof_kinds(void);
int coopcap_prof_enable(int on, coopcap_stream_t stream);
int coopcap_prof_report(double* ms, double* flops, double* bytes, long long* launches, int nkinds);

/* SM clock measured on the device: one warp spins for ~spin_ns and compares clock64 with
 * globaltimer; writes MHz to *mhz_out (device float).  Costs ~spin_ns of stream time and no driver
 * query, so it can be dropped between steps of a timed region. */
int coopcap_measure_sm_clock(float* mhz_out, int spin_ns, coopcap_stream_t stream);

/* sizeof() of the structs above, for binding self-checks: which = 0 gemm_args, 1 speaker_pack,
 * 2 speaker, 3 speaker_grads, 4 listener_pack, 5 listener, 6 listener_grads, 7 cider, 8 beam. */
int coopcap_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* COOPCAP_H_ */
