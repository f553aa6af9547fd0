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

#ifdef __cplusplus
}
#endif
#endif /* COOPCAP_H_ */
