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
  /* Optional grouped row addressing (0 = plain row*ld): element offset of row r is
   *   (r / group) * group_stride + (r % group) * ld.
   * C (and aux) use c_group; residual uses r_group (r_group_stride = 0 makes the residual a
   * periodic [r_group, N] table, used for the patch-embedding position+temporal add). */
  int64_t c_group, c_group_stride, r_group, r_group_stride;
  int32_t block_n;    /* 0 = auto, else 128 or 256 */
  int32_t max_ctas;   /* 0 = one persistent CTA per SM */
  int32_t cta_pair;   /* 0 = auto (2-CTA cta_group::2 pairs on 256x256 tiles when M, N >= 256), 1 = never, 2 = force */
  int32_t reserved;
} XpGemm;

int xp_gemm(const XpGemm* g, void* stream);

/* ------------------------------------------------------------- row kernels --
 * Row addressing shared by the row-wise kernels: element offset of logical row r is
 *   offsets[r]                                            if offsets != NULL (device i64 array)
 *   (r / group) * group_stride + (r % group) * ld         if group > 0
 *   r * ld                                                otherwise.
 * This is how the kernels skip the M global tokens of each video, pick the CLS row
 * (CLIP_ViP.py:891) or the EOS row (CLIP_ViP.py:776) without a gather copy. */
typedef struct XpRowMap {
  int64_t group, group_stride, ld;
  const int64_t* offsets;
} XpRowMap;

/* nn.LayerNorm forward (CLIP_ViP.py:447,458,881,892,771): bf16 in/out, fp32 gamma/beta/statistics. */
int xp_layernorm_fwd(const void* x, const XpRowMap* xmap, void* y, const XpRowMap* ymap, const float* gamma,
                     const float* beta, float* mean, float* rstd, int64_t rows, int32_t C, float eps, void* stream);
/* LayerNorm fused with the residual add of the block it opens (CLIPEncoderLayer, CLIP_ViP.py:445-460: `hidden = residual +
 * branch; hidden = layer_normN(hidden)`), with the residual stream kept in fp32 as under the reference's autocast:
 *   s = x (+ add_bf16);  sum_out = s (optional);  y = LayerNorm(s).
 * x is bf16, fp32 or fp16 (x_dtype) and so is y (y_dtype); sum_out is fp32, or fp16 when x is fp16 (saturating; the
 * reference's own training precision under apex O2, run_pretrain.py:234-236).  add / sum_out / their maps may be NULL. */
int xp_layernorm_add_fwd(const void* x, const XpRowMap* xmap, int32_t x_dtype, const void* add_bf16, const XpRowMap* addmap,
                         void* sum_out, const XpRowMap* summap, void* y, const XpRowMap* ymap, int32_t y_dtype,
                         const float* gamma, const float* beta, float* mean, float* rstd, int64_t rows, int32_t C, float eps,
                         void* stream);
/* LayerNorm backward; dx = LN'(dy) + dres (the residual-branch gradient, may be NULL);
 * dgamma/dbeta are ACCUMULATED (fp32 atomics) so they can point at .grad buffers.  dres_colsum (optional, needs dres):
 * ACCUMULATES sum_rows dres — the bias gradient of the Linear that closes the other branch of that residual add
 * (fc2.bias / out_proj.bias of CLIPEncoderLayer, CLIP_ViP.py:445-460), so no separate column-sum pass reads dres again.
 * x (the saved LayerNorm input) is bf16 or fp32 (x_dtype); dy, dres, dx are bf16. */
int xp_layernorm_bwd(const void* dy, const XpRowMap* dymap, const void* x, const XpRowMap* xmap, int32_t x_dtype,
                     const float* gamma, const float* mean, const float* rstd, const void* dres, const XpRowMap* drmap, void* dx,
                     const XpRowMap* dxmap, float* dgamma, float* dbeta, float* dres_colsum, int64_t rows, int32_t C,
                     void* stream);
/* x / x.norm(dim=-1, keepdim=True) (CLIP_ViP.py:1148-1149), fp32. */
int xp_l2norm_fwd(const float* x, float* y, float* inv_norm, int32_t rows, int32_t C, void* stream);
int xp_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, void* dx_bf16, int32_t rows, int32_t C,
                  float scale, void* stream);
/* out[c] += scale * sum_r x[r,c]: bias gradients of every nn.Linear. */
int xp_colsum_bf16(const void* x, int64_t ld, float* out, int64_t rows, int32_t C, float scale, void* stream);
/* fp32 master parameter -> bf16 compute copy. */
int xp_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------- embeddings --*/
enum { XP_DTYPE_F32 = 0, XP_DTYPE_BF16 = 1, XP_DTYPE_F16 = 2 };

/* im2col of nn.Conv2d(3, width, kernel=stride=patch, bias=False) (CLIP_ViP.py:157-159,178-179):
 * video [frames,3,H,W] -> patches bf16 [frames*(H/p)*(W/p), 3*p*p]; the conv itself then runs as xp_gemm. */
int xp_vip_patchify(const void* video, int32_t dtype, void* patches_bf16, int64_t frames, int32_t H, int32_t W,
                    int32_t patch, void* stream);
/* The reference's input transform fused into the patch extraction: frames_hwc uint8 [frames, H, W, 3] as
 * the decoder delivers them -> `.permute(0,3,1,2).float() / 255.` (CLIP-ViP/src/datasets/dataset_pretrain_stage1_all_source.py:182)
 * -> Normalize(mean, std) (init_transform_dict_simple, CLIP-ViP/src/datasets/dataloader.py:209-233; Resize / CenterCrop are
 * the identity at the input resolution) -> the bf16 patch matrix of xp_vip_patchify.  IEEE fp32 arithmetic, one rounding to
 * bf16: bit-identical to casting the reference's fp32 tensor.  mean3 / std3 are HOST arrays of 3 floats. */
int xp_vip_patchify_u8(const uint8_t* frames_hwc, void* patches_bf16, int64_t frames, int32_t H, int32_t W, int32_t patch,
                       const float* mean3, const float* std3, void* stream);
/* CLIP_ViP.py:170-176,183-195: table[t*L+l] = interp(temporal_embedding)[t] + position_embedding[1+l] (bf16,
 * [T*L, C]) and the M = 1 + add_cls_num global rows x[b, m] = (class_embedding | added_cls[m-1]) + position_embedding[0]
 * written into x_bf16 [B, M+T*L, C].  temporal may be NULL (if_use_temporal_embed = 0). */
int xp_vip_embed_tables(const float* pos, const float* temporal, const float* cls, const float* added,
                        void* table_bf16, void* x_bf16, int32_t B, int32_t T, int32_t L, int32_t M, int32_t C,
                        int32_t temporal_size, void* stream);
/* Backward of the above: d_patch bf16 [B, T*L, C] and d_global bf16 [B, M, C] (the two compact halves of
 * d_embeddings) accumulated (fp32) into the four parameter gradients. */
int xp_vip_embed_bwd(const void* d_patch_bf16, const void* d_global_bf16, float* d_pos, float* d_temporal, float* d_cls,
                     float* d_added, int32_t B, int32_t T, int32_t L, int32_t M, int32_t C, int32_t temporal_size,
                     void* stream);
/* CLIPTextEmbeddings.forward (CLIP_ViP.py:222-225): x[r] = token_embedding[ids[r]] + position_embedding[r % Lt].
 * ids are int64 and indexed bit-exactly; *err_flag is set to 1 if any id is outside [0, vocab). */
int xp_text_embed_fwd(const int64_t* ids, const float* tok, const float* pos, void* x_bf16, int32_t rows, int32_t Lt,
                      int32_t C, int32_t vocab, int32_t* err_flag, void* stream);
int xp_text_embed_bwd(const int64_t* ids, const void* dx_bf16, float* d_tok, float* d_pos, int32_t rows, int32_t Lt,
                      int32_t C, int32_t vocab, void* stream);
/* EOS pooling row (CLIP_ViP.py:776): offsets[b] = (b*Lt + first argmax_s ids[b,s]) * C; index[b] optional. */
int xp_eos_offsets(const int64_t* ids, int64_t* offsets, int32_t* index, int32_t B, int32_t Lt, int32_t C,
                   void* stream);

/* ------------------------------------------------------------ ViP attention --
 * CLIPAttention.forward2 (CLIP_ViP.py:332-381) between the QKV projection and out_proj.
 *   qkv  bf16 [B*S, 3C]  columns [q | k | v], head h at [h*64, h*64+64), q already scaled by 64**-0.5
 *   out  bf16 [B*S, C]   (the tensor out_proj consumes; rows ordered [M global, frame0 L, frame1 L, ...])
 *   lse  f32  [B, H, S]  log-sum-exp of every query row (saved for backward)
 * S = M + T*L, head_dim 64, M + L <= 208.  workspace: xp_vip_attention_workspace_bytes() bytes. */
int64_t xp_vip_attention_workspace_bytes(int32_t B, int32_t H, int32_t T, int32_t M);
int xp_vip_attention_fwd(const void* qkv, void* out, float* lse, float* workspace, int32_t B, int32_t H, int32_t T,
                         int32_t L, int32_t M, int32_t C, void* stream);
/* Same contract, tcgen05/TMEM kernel (S and P·V on the 5th-gen tensor cores, softmax thread-per-TMEM-lane). */
int xp_vip_attention_fwd_tc(const void* qkv, void* out, float* lse, float* workspace, int32_t B, int32_t H, int32_t T,
                            int32_t L, int32_t M, int32_t C, void* stream);
/* First half of the above (frame rows + per-frame partials of the global rows, before the combine step). */
int xp_vip_attention_fwd_tc_partial(const void* qkv, void* out, float* lse, float* workspace, int32_t B, int32_t H,
                                    int32_t T, int32_t L, int32_t M, int32_t C, void* stream);
/* dqkv bf16 [B*S, 3C] = gradient w.r.t. the (un-scaled-q) projection outputs, i.e. the dq part already carries
 * q_scale (CLIP_ViP.py:341), so the QKV dgrad/wgrad GEMMs treat the three thirds uniformly. */
int xp_vip_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                         float* workspace, int32_t B, int32_t H, int32_t T, int32_t L, int32_t M, int32_t C,
                         float q_scale, void* stream);

/* tcgen05/TMEM backward: S, dP, dV, dK, dQ on the 5th-gen tensor cores.  delta f32 [B, H, S] is scratch that
 * receives rowsum(dout * out). */
int xp_vip_attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                            float* workspace, float* delta, int32_t B, int32_t H, int32_t T, int32_t L, int32_t M,
                            int32_t C, float q_scale, void* stream);
int xp_vip_attention_bwd_tc_partial(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                    float* workspace, float* delta, int32_t B, int32_t H, int32_t T, int32_t L,
                                    int32_t M, int32_t C, float q_scale, void* stream);

/* ------------------------------------------------------- text-tower attention --
 * CLIPAttention.forward (CLIP_ViP.py:266-330) between the QKV projection and out_proj, with the causal mask
 * (CLIP_ViP.py:788-797) and the padding mask built from attention_mask int64 [B, Lt] (CLIP_ViP.py:50-61).
 * qkv bf16 [B*Lt, 3C] (q pre-scaled), out bf16 [B*Lt, C], probs f32 [B, H, Lt, Lt] (saved for backward). Lt <= 96. */
int xp_text_attention_fwd(const void* qkv, const int64_t* mask, void* out, float* probs, int32_t B, int32_t H,
                          int32_t Lt, int32_t C, void* stream);
int xp_text_attention_bwd(const void* qkv, const void* dout, const float* probs, void* dqkv, int32_t B, int32_t H,
                          int32_t Lt, int32_t C, float q_scale, void* stream);

/* ------------------------------------------------------------------ InfoNCE --
 * NCELearnableTempLoss.forward (loss.py:134-141) on the gathered [N, d] fp32 embeddings.  The logits GEMM and the
 * two gradient GEMMs run through xp_gemm; these are the pieces around them:
 *   xp_nce_split:        x f32 [rows, d] -> x3 bf16 [rows, 3d] = [hi|hi|lo] (pattern 0) or [hi|lo|hi] (pattern 1),
 *                        so that x3_a . x3_b^T = hi*hi + hi*lo + lo*hi (fp32-grade logits on bf16 tensor cores);
 *                        hi_bf16 (optional) receives the plain bf16 copy used by the gradient GEMMs.
 *   xp_nce_softmax_grad: z f32 [N, N] (row pitch ld, shared with g_scaled) = V T^T (unscaled) -> row/col LSE of exp(logit_scale)*z, the scalar loss
 *                        (overwritten), d_logit_scale (ACCUMULATED) and g_scaled bf16 [N,N] = exp(logit_scale) * dL/dZ. */
int xp_nce_split(const float* x, void* x3_bf16, void* hi_bf16, int32_t rows, int32_t d, int32_t pattern, void* stream);
int xp_nce_softmax_grad(const float* z, const float* logit_scale, float* lse_rows, float* lse_cols, void* g_scaled_bf16,
                        float* loss, float* d_logit_scale, int32_t N, int64_t ld, void* stream);
/* Fused exchange + loss: replaces `hvd.allgather(vis)`, `hvd.allgather(txt)` (CLIP-ViP/src/pretrain/run_pretrain.py:344-345;
 * rank-major concat, semantics pinned by LF-VILA/src/utils/dist.py:21-41) AND NCELearnableTempLoss.forward (loss.py:134-141)
 * with ONE cooperative kernel (csrc/nce_fused.cu): device-side flag barrier over peer-mapped exchange buffers, logits tiles
 * on tcgen05 whose operand rows are loaded straight from the owning peer's memory over NVLink (hi/lo split in the producer),
 * row/column log-sum-exps, loss (overwritten), d logit_scale (overwritten), g_scaled bf16 [N, ld_g] = exp(logit_scale)*dL/dZ,
 * and the bf16 copies vis_hi / txt_hi [N, d] that the local gradient GEMMs use.  N = world * b <= 1536.
 *   mode 0: peer_bufs = device array of `world` exchange-buffer base pointers (own buffer at [rank]); every buffer is
 *           xp_nce_gather_exchange_bytes() large, zero-initialised once, and mapped by all ranks (symmetric memory);
 *           vis_local / txt_local fp32 [b, d] are published by the kernel; `epoch` must increase by 1 per call (from 1).
 *   mode 1: no exchange: peer_bufs = device array of 2*world pointers, [r] = rank r's vis rows, [world + r] = its txt rows
 *           (fp32 [b, d]) in local memory (single process, or rows pre-gathered by another transport).
 * workspace: xp_nce_gather_workspace_bytes(N) bytes, zero-initialised once (it holds the kernel's barrier counters). */
typedef struct XpNceGather {
  const float* vis_local;
  const float* txt_local;
  void* const* peer_bufs;
  const float* logit_scale;
  void* g_scaled;
  void* vis_hi;
  void* txt_hi;
  float* loss;
  float* d_logit_scale;
  float* workspace;
  int32_t rank, world, b, d;
  uint32_t epoch;
  int32_t mode;
  int64_t ld_g;
} XpNceGather;
int64_t xp_nce_gather_exchange_bytes(int32_t b, int32_t d, int32_t world);
int64_t xp_nce_gather_workspace_bytes(int32_t N);
int xp_nce_gather_fused(const XpNceGather* args, void* stream);
/* NCELearnableTempLoss_vsc_fc.forward, CLIP-ViP/src/optimization/loss.py:288-324 (the released pre-training default:
 * video x subtitle, video x caption, frame x caption): za = V T^T, zb = V C^T, zd = I C^T, fp32 [N, N] with row pitch ld,
 * unscaled.  Writes the scalar loss (overwritten), ACCUMULATES d_logit_scale, and the three gradient matrices
 * g* = exp(logit_scale) * dL/d(s z*) as bf16 [N, N] (row pitch ld).  stats: 6*N floats of scratch. */
int xp_nce_vsc_fc(const float* za, const float* zb, const float* zd, const float* logit_scale, float* stats, void* ga_bf16,
                  void* gb_bf16, void* gd_bf16, float* loss, float* d_logit_scale, int32_t N, int64_t ld, void* stream);

/* ---- BASELINE.json config #4: HD-VILA TimeSformer (divided space-time attention), hd-vila/src/modeling/timesformer.py
 *
 * xp_seg_attention_{fwd,bwd}: multi-head attention (head_dim 64) over strided "sequences" of a token-major fused
 * [n_rows, ld_qkv] bf16 buffer (columns [q|k|v], head h at h*64; q pre-scaled by head_dim**-0.5).  Replaces
 * Attention.forward (timesformer.py:156-173) TOGETHER WITH the einops rearranges around it (Block.forward :210-219):
 *   sequence s starts at row (s / inner) * outer_stride + (s % inner) * inner_stride, its tokens are tok_stride rows
 *   apart, it has min(seq_len, rows that fit) tokens, and token i attends to token j iff i / seg_len == j / seg_len
 *   (seg_len >= seq_len: dense).
 *   temporal ('(b h w) t m'):  G = 64 / T groups per sequence: seq_len = G*T, seg_len = T, inner = 1,
 *                              outer_stride = G*T, tok_stride = 1, n_seq = ceil(n_rows / (G*T))
 *   spatial  ('(b t) (h w) m'): seq_len = seg_len = H*W, inner = T, outer_stride = H*W*T, inner_stride = 1,
 *                              tok_stride = T, n_seq = B*T
 * out: [n_rows, ld_out] bf16 (same row order as qkv); lse, delta: [heads, n_rows] fp32 (delta is scratch written by bwd);
 * dqkv: [n_rows, ld_qkv] bf16, every (row, head) slice covered by a sequence is overwritten; dq is multiplied by q_scale. */
typedef struct XpSegAttn {
  int64_t n_rows;
  int64_t ld_qkv, ld_out;
  int64_t outer_stride, inner_stride, tok_stride;
  int32_t heads, n_seq, seq_len, seg_len, inner, reserved;
  /* Window-attention extensions — LF-VILA WindowAttention3D.forward, LF-VILA/src/models/video_encoder.py:135-164, together
   * with the pad / roll / window_partition / window_reverse / crop around it (:214-243).  All optional (NULL / 0):
   *   row_index    int32 [n_seq, seq_len]: token row of every window position (replaces the stride pattern; dense sequences).
   *                Built once per feature-map shape by applying the reference's own pad/roll/partition to an index tensor;
   *                zero-padded positions point at extra all-zero input rows, whose k, v are the qkv bias exactly as in the
   *                reference.
   *   bias         fp32 [bias_windows, heads, seq_len, seq_len] added to the logits before the softmax: relative-position
   *                bias (+ the 0 / -100 shift mask of window type s % bias_windows)
   *   ds_out       (bwd) bf16 [n_seq, heads, seq_len, seq_len]: dL/dlogits, whose sum over windows is the bias gradient
   *   head_dim     64 (or 0) or 32 */
  const int32_t* row_index;
  const float* bias;
  void* ds_out;
  int32_t bias_windows;
  int32_t head_dim;
} XpSegAttn;
int xp_seg_attention_fwd(const void* qkv, void* out, float* lse, const XpSegAttn* desc, void* stream);
int xp_seg_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                         const XpSegAttn* desc, float q_scale, void* stream);

/* Token assembly, TimeSformer.forward timesformer.py:481-509: x [B,T,C,H*W] (XP_DTYPE_*) -> tokens bf16 [(b, p, t), C]
 * (rows in the reference's (h w t) order) = x[b,t,:,p] + pos[p,:] + time[t,:]; pos [H*W, C] / time [T, C] fp32 are the
 * (already interpolated) tables, NULL = no table (plain tokenisation, used for the output gradient).
 * xp_tsf_untokenize is the inverse layout change (tokens -> [B,T,C,H*W]): the module output (:523) and d(x). */
int xp_tsf_embed_fwd(const void* x, int32_t x_dtype, const float* pos, const float* time, void* tokens_bf16, int32_t B,
                     int32_t T, int32_t C, int32_t HW, void* stream);
int xp_tsf_untokenize(const void* tokens_bf16, void* x, int32_t x_dtype, int32_t B, int32_t T, int32_t C, int32_t HW,
                      void* stream);
/* Stochastic depth on a residual branch (DropPath, timesformer.py:98-121, as Block.forward applies it :212,:218,:225):
 * out[r,:] = (residual ? residual[r,:] : 0) + scale[r] * x[r,:], all [rows, C] bf16 contiguous, scale fp32 per row
 * (0 or 1/keep_prob of the row's sample/group).  out may alias x. */
int xp_rowscale_bf16(const void* x, const float* scale, const void* residual, void* out, int64_t rows, int32_t C,
                     void* stream);

/* ---- BASELINE.json config #5 helpers (LF-VILA Swin-3D, LF-VILA/src/models/video_encoder.py)
 * xp_layernorm_wide_*: nn.LayerNorm over 1024 < C <= 4096 contiguous columns — PatchMerging.norm (4C = 2048, :281,:304).
 * xp_gather_rows_bf16 / xp_scatter_rows_bf16: out[i, :] = src[index[i], :] (zeros for index < 0) and its inverse
 * dst[index[i], :] = in[i, :] — the 2x2 neighbour concatenation of PatchMerging.forward (:292-301; out viewed as
 * [n/4... , 4C]) with its odd-size zero padding, and its backward. */
int xp_layernorm_wide_fwd(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                          int64_t rows, int32_t C, float eps, void* stream);
int xp_layernorm_wide_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                          float* dgamma, float* dbeta, int64_t rows, int32_t C, void* stream);
int xp_gather_rows_bf16(const void* src, const int32_t* index, void* out, int64_t n_items, int32_t C, void* stream);
int xp_scatter_rows_bf16(const void* in, const int32_t* index, void* dst, int64_t n_items, int32_t C, void* stream);

/* ---- retrieval evaluation.  Replaces the numpy calls of validate() (CLIP-ViP/src/pretrain/
 * run_pretrain.py:173-176, tasks/run_video_retrieval.py:155-172) on CLIP-ViP/src/utils/metrics.py:
 *   xp_sim_f32      cal_cossim (:3-5): out[Na, Nb] (row pitch ld) = a[Na, d] b[Nb, d]^T, fp32 FFMA accumulation
 *   xp_dsl_reweight sim *= softmax(theta * sim, axis=0) in place (np_softmax :7-39 as used at run_video_retrieval.py:169-170);
 *                   col_scratch: 2*cols floats
 *   xp_rank_counts  compute_metrics' rank search (:41-48) without the sort: greater[i] / equal[i] = number of entries of row i
 *                   (transpose != 0: column i) strictly larger than / equal to sim[i, i]; the reference's rank list is
 *                   {greater[i] + t : 0 <= t < equal[i]} (ties counted once per tied entry, as np.where(ind == 0) does). */
int xp_sim_f32(const float* a, const float* b, float* out, int32_t Na, int32_t Nb, int32_t d, int64_t ld, void* stream);
int xp_dsl_reweight(float* sim, int32_t rows, int32_t cols, int64_t ld, float theta, float* col_scratch, void* stream);
int xp_rank_counts(const float* sim, int32_t N, int64_t ld, int32_t transpose, int32_t* greater, int32_t* equal, void* stream);

/* ---- the optimizer step.  Replaces AdamW.step (CLIP-ViP/src/optimization/adamw.py:40-103) and
 * torch.nn.utils.clip_grad_norm_ as called at pretrain/run_pretrain.py:408-411,422 — one table-driven launch over all
 * parameters instead of ~10 elementwise launches per parameter.
 *
 * table_dev:     device array of XpOptTensor, one per parameter (all fp32, contiguous).  step_size =
 *                lr * sqrt(1 - beta2^t) / (1 - beta1^t) (or lr when correct_bias is off) and decay = lr * weight_decay are
 *                computed by the host per parameter group; p_bf16 (optional) receives the bf16 copy of the updated p.
 * block_map_dev: device array of n_blocks {tensor index, chunk index} int32 pairs; chunk c of a tensor covers elements
 *                [c * xp_opt_chunk_elems(), ...).  Built once per parameter set by the host.
 * xp_opt_grad_norm:  norm_out_dev[0] = 2-norm over every g in the table, norm_out_dev[1] = min(1, max_norm/(norm+1e-6))
 *                    (1 if max_norm <= 0); partial_dev is n_blocks floats of scratch.  Deterministic.
 * xp_opt_scale_grads: g *= norm_dev[1] in place (clip_grad_norm_ used on its own).
 * xp_opt_adamw_step:  the fused update; norm_dev (optional) = the pair above, its coefficient is applied to g on the fly. */
typedef struct XpOptTensor {
  void* p;
  const void* g;
  void* m;
  void* v;
  void* p_bf16;
  int64_t n;
  float step_size;
  float decay;
  int32_t reserved[2];
} XpOptTensor;
int32_t xp_opt_chunk_elems(void);
int xp_opt_grad_norm(const XpOptTensor* table_dev, const int32_t* block_map_dev, int32_t n_blocks, float* partial_dev,
                     float max_norm, float* norm_out_dev, void* stream);
int xp_opt_scale_grads(const XpOptTensor* table_dev, const int32_t* block_map_dev, int32_t n_blocks, const float* norm_dev,
                       void* stream);
int xp_opt_adamw_step(const XpOptTensor* table_dev, const int32_t* block_map_dev, int32_t n_blocks, const float* norm_dev,
                      float beta1, float beta2, float eps, void* stream);
/* Multi-tensor refresh of the bf16 compute copies (replaces the reference's implicit "the module reads its own fp32
 * parameters", CLIP_ViP.py:445-460): per row, g = fp32 source, p_bf16 = bf16 destination, or if that is null p = fp32
 * destination (plain copy); n elements.  One launch for all weights of a tower, run on every forward. */
int xp_cast_table(const XpOptTensor* table_dev, const int32_t* block_map_dev, int32_t n_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XPRETRAIN_B200_H */
