/*
 * xpretrain_b200 — C ABI of the B200-native CLIP-ViP / HD-VILA hot path.
 *
 * The reference (microsoft/XPretrain) is 100 % Python: it has no FFI or
 * plugin boundary for this path, so the seam a maintainer binds is the set of
 * torch ops its nn.Modules call.  Each entry point below names the reference
 * lines it replaces.  Conventions:
 *   - every pointer is a raw DEVICE pointer owned by the caller (PyTorch
 *     allocates; nothing here allocates or frees device memory);
 *   - `stream` is a cudaStream_t passed as void*;
 *   - bf16 = __nv_bfloat16 bits, f32 = IEEE float, i64 = int64_t;
 *   - return 0 on success, negative on error; xp_last_error() gives the text.
 *     There is no CPU fallback: without a B200 every call fails loudly.
 */
#ifndef XPRETRAIN_B200_H
#define XPRETRAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XP_ABI_VERSION 1

int xp_version(void);
const char* xp_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t xp_launch_count(void);
void xp_launch_count_reset(void);

/* ------------------------------------------------------------------ GEMM --
 * C[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )      (tcgen05 / TMEM / TMA)
 * Replaces every nn.Linear on the path and its autograd:
 *   forward  y = x W^T + b         CLIP_ViP.py:341-343,379 (q/k/v/out_proj), :393-395 (fc1/fc2),
 *                                   :1141-1145 (visual/text projection), :178 (patch conv as im2col GEMM)
 *   dgrad    dx = dy W             a_layout=0, b_layout=1
 *   wgrad    dW += dy^T x          a_layout=1, b_layout=1, out=XP_OUT_F32_ATOMIC
 * a_layout: 0 = A stored [M,K] (K contiguous), 1 = A stored [K,M] (M contiguous)
 * b_layout: 0 = B stored [N,K] (K contiguous, nn.Linear.weight), 1 = B stored [K,N]
 * Epilogue, in this order:  v = alpha*acc; v += bias[n]; if n < scale_cols: v *= col_scale
 *   (CLIP_ViP.py:341 scales q AFTER the bias); act; v += residual[m,n]; store.
 */
enum { XP_ACT_NONE = 0, XP_ACT_QUICK_GELU = 1, XP_ACT_DQUICK_GELU = 2, XP_ACT_GELU_ERF = 3, XP_ACT_DGELU_ERF = 4 };
enum { XP_OUT_BF16 = 0, XP_OUT_F32 = 1, XP_OUT_F32_ATOMIC = 2 };

typedef struct XpGemm {
  const void* a;        /* bf16 */
  const void* b;        /* bf16 */
  void* c;              /* bf16 or f32 per `out` */
  const float* bias;    /* f32 [N] or NULL */
  const void* residual; /* bf16 [M, ldr] or NULL */
  void* aux;            /* bf16 [M, ld_aux]: QUICK_GELU/GELU_ERF store the pre-activation here (may be NULL);
                           DQUICK_GELU/DGELU_ERF read the pre-activation from here */
  int64_t M, N, K;
  int64_t lda, ldb, ldc, ldr, ld_aux; /* leading dimensions in elements */
  int32_t a_layout, b_layout;
  int32_t act, out;
  int32_t splits;     /* split-K factor (>1 only with XP_OUT_F32_ATOMIC) */
  int32_t scale_cols; /* columns [0, scale_cols) are multiplied by col_scale */
  float alpha, col_scale;
  int32_t block_n;    /* 0 = auto, else 128 or 256 */
  int32_t max_ctas;   /* 0 = one persistent CTA per SM */
} XpGemm;

int xp_gemm(const XpGemm* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XPRETRAIN_B200_H */
