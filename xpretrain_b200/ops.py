"""Tensor-level wrappers over the C ABI.  PyTorch supplies device memory and the current stream only;
every computation below runs in the hand-written sm_100a kernels of libxpretrain_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import XpGemm, XpRowMap, XpSegAttn, check, lib

bf16, f32 = torch.bfloat16, torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.XpError("xpretrain_b200 kernels need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def launch_count() -> int:
    return int(lib().xp_launch_count())


def reset_launch_count() -> None:
    lib().xp_launch_count_reset()


# ------------------------------------------------------------------------------------------ GEMM
def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int, ldc: int,
         a_layout: int = 0, b_layout: int = 0, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, ldr: int = 0, aux: Optional[torch.Tensor] = None, ld_aux: int = 0,
         act: int = _lib.ACT_NONE, out_mode: int = _lib.OUT_BF16, splits: int = 1, scale_cols: int = 0,
         col_scale: float = 1.0, alpha: float = 1.0, c_group: int = 0, c_group_stride: int = 0, r_group: int = 0,
         r_group_stride: int = 0, block_n: int = 0, a_offset: int = 0, b_offset: int = 0, c_offset: int = 0,
         cta_pair: int = 0) -> None:
    """out = epilogue(alpha * A @ B^T); offsets are in elements from the tensors' data pointers."""
    assert a.dtype == bf16 and b.dtype == bf16
    if bias is not None:
        assert bias.dtype == f32 and bias.is_contiguous()
    g = XpGemm()
    g.a = _p(a) + a_offset * 2
    g.b = _p(b) + b_offset * 2
    g.c = _p(out) + c_offset * out.element_size()
    g.bias = _p(bias)
    g.residual = _p(residual)
    g.aux = _p(aux)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc, g.ldr, g.ld_aux = lda, ldb, ldc, ldr, ld_aux
    g.a_layout, g.b_layout, g.act, g.out = a_layout, b_layout, act, out_mode
    g.splits, g.scale_cols, g.alpha, g.col_scale = splits, scale_cols, alpha, col_scale
    g.c_group, g.c_group_stride, g.r_group, g.r_group_stride = c_group, c_group_stride, r_group, r_group_stride
    g.block_n, g.max_ctas, g.cta_pair = block_n, _sm_limit, cta_pair
    if _gemm_timer is None:
        check(lib().xp_gemm(C.byref(g), _stream()), "xp_gemm")
    else:  # bench.py's roofline leg: CUDA events on the launching stream around this launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib().xp_gemm(C.byref(g), _stream()), "xp_gemm")
        e1.record()
        _gemm_timer.append((2.0 * M * N * K, e0, e1))


_gemm_timer = None
_sm_limit = 0


def set_sm_limit(n: int) -> None:
    """Cap the persistent GEMM grids at n CTAs (0 = all SMs).  A data-parallel job reserves a few SMs this way for the NCCL
    kernels of the overlapped gradient all-reduce (NCCL_MAX_CTAS), so that they never displace a persistent GEMM CTA — whose
    tiles would then run as a second, nearly empty wave (`gemm_ms_per_step` 71.3 -> 74.7 ms from 1 to 8 GPUs)."""
    global _sm_limit
    _sm_limit = max(0, int(n)) // 2 * 2       # CTA pairs: keep it even


def set_gemm_timer(records) -> None:
    """records: a list receiving (flops, start_event, end_event) per GEMM launch, or None to switch timing off."""
    global _gemm_timer
    _gemm_timer = records


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, **kw) -> None:
    """out[M,N] = x[M,K] @ w[N,K]^T (+bias, epilogue)."""
    M, K = x.shape
    N = w.shape[0]
    gemm(x, w, out, M=M, N=N, K=K, lda=x.stride(0), ldb=w.stride(0), ldc=out.stride(0), bias=bias, **kw)


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, dx: torch.Tensor, **kw) -> None:
    """dx[M,K] = dy[M,N] @ w[N,K]   (w in its nn.Linear layout: MN-major B operand)."""
    M, N = dy.shape
    K = w.shape[1]
    gemm(dy, w, dx, M=M, N=K, K=N, lda=dy.stride(0), ldb=w.stride(0), ldc=dx.stride(0), b_layout=1, **kw)


def wgrad_plan(n_out: int, n_in: int, rows: int, sms: int = 0):
    """(block_n, splits) for a weight-gradient GEMM: the split-K factor that fills whole waves of the persistent grid.
    Outputs of at least 256 x 256 run on CTA pairs (256 x 256 tiles, sms/2 clusters), smaller ones on single CTAs."""
    sms = sms or _sm_limit or 148
    pair = n_out >= 256 and n_in >= 256
    bn = 256 if n_in >= 256 else 128
    if pair:
        tiles, slots = ((n_out + 255) // 256) * ((n_in + 255) // 256), sms // 2
    else:
        tiles, slots = ((n_out + 127) // 128) * ((n_in + bn - 1) // bn), sms
    best, best_eff = 1, 0.0
    for s in range(1, 33):
        if s > 1 and rows // s < 1024:
            break
        total = tiles * s
        eff = total / (((total + slots - 1) // slots) * slots)
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
    return bn, best


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, alpha: float = 1.0) -> None:
    """dw[N,K] += dy[rows,N]^T @ x[rows,K]  (fp32 atomics; both operands MN-major, split-K over the rows)."""
    rows, N = dy.shape
    K = x.shape[1]
    assert dw.dtype == f32
    bn, splits = wgrad_plan(N, K, rows)
    gemm(dy, x, dw, M=N, N=K, K=rows, lda=dy.stride(0), ldb=x.stride(0), ldc=dw.stride(0), a_layout=1, b_layout=1,
         out_mode=_lib.OUT_F32_ATOMIC, splits=splits, block_n=bn, alpha=alpha)


# ------------------------------------------------------------------------------------ row kernels
def rowmap(ld: int, group: int = 0, group_stride: int = 0, offsets: Optional[torch.Tensor] = None) -> XpRowMap:
    m = XpRowMap()
    m.group, m.group_stride, m.ld, m.offsets = group, group_stride, ld, _p(offsets)
    return m


def layernorm_fwd(x, xmap, y, ymap, gamma, beta, mean, rstd, rows: int, C_: int, eps: float, x_off=0, y_off=0,
                  add=None, addmap=None, add_off=0, sum_out=None, summap=None):
    """y = LayerNorm(x (+ add)); offsets in ELEMENTS of the respective tensor.  x / y may be bf16 or fp32 (the fp32 residual
    stream); `add` is the bf16 branch output folded in before the normalisation, `sum_out` (fp32) receives x + add."""
    dt = {torch.bfloat16: _lib.DTYPE_BF16, torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16}
    assert add is None or add.dtype == bf16
    assert sum_out is None or sum_out.dtype == (torch.float16 if x.dtype == torch.float16 else f32)
    check(lib().xp_layernorm_add_fwd(_p(x) + x_off * x.element_size(), C.byref(xmap), dt[x.dtype],
                                     (_p(add) + add_off * 2) if add is not None else None,
                                     C.byref(addmap) if addmap is not None else None, _p(sum_out),
                                     C.byref(summap) if summap is not None else None, _p(y) + y_off * y.element_size(),
                                     C.byref(ymap), dt[y.dtype], _p(gamma), _p(beta), _p(mean), _p(rstd), rows, C_, eps,
                                     _stream()), "xp_layernorm_add_fwd")


def layernorm_bwd(dy, dymap, x, xmap, gamma, mean, rstd, dres, drmap, dx, dxmap, dgamma, dbeta, rows: int, C_: int,
                  dy_off=0, x_off=0, dres_off=0, dx_off=0, dres_colsum=None):
    """x: the saved LayerNorm input, bf16, fp32 or fp16; dy / dres / dx bf16.  Offsets in elements."""
    dt = {torch.bfloat16: _lib.DTYPE_BF16, torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16}
    check(lib().xp_layernorm_bwd(_p(dy) + dy_off * 2, C.byref(dymap), _p(x) + x_off * x.element_size(), C.byref(xmap),
                                 dt[x.dtype], _p(gamma), _p(mean), _p(rstd),
                                 (_p(dres) + dres_off * 2) if dres is not None else None,
                                 C.byref(drmap) if drmap is not None else None, _p(dx) + dx_off * 2, C.byref(dxmap),
                                 _p(dgamma), _p(dbeta), _p(dres_colsum), rows, C_, _stream()), "xp_layernorm_bwd")


def l2norm_fwd(x, y, inv_norm):
    rows, C_ = x.shape
    check(lib().xp_l2norm_fwd(_p(x), _p(y), _p(inv_norm), rows, C_, _stream()), "xp_l2norm_fwd")


def l2norm_bwd(dy, y, inv_norm, dx_bf16, scale: float = 1.0):
    rows, C_ = y.shape
    check(lib().xp_l2norm_bwd(_p(dy), _p(y), _p(inv_norm), _p(dx_bf16), rows, C_, scale, _stream()), "xp_l2norm_bwd")


def colsum(x: torch.Tensor, out: torch.Tensor, scale: float = 1.0):
    rows, C_ = x.shape
    check(lib().xp_colsum_bf16(_p(x), x.stride(0), _p(out), rows, C_, scale, _stream()), "xp_colsum_bf16")


def cast_bf16(src: torch.Tensor, dst: torch.Tensor, dst_offset: int = 0):
    assert src.dtype == f32 and dst.dtype == bf16 and src.is_contiguous()
    check(lib().xp_cast_f32_bf16(_p(src), _p(dst) + dst_offset * 2, src.numel(), _stream()), "xp_cast_f32_bf16")


# ------------------------------------------------------------------------------------- embeddings
_DT = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def vip_patchify(video: torch.Tensor, patches: torch.Tensor, patch: int):
    frames = video.numel() // (3 * video.shape[-2] * video.shape[-1])
    check(lib().xp_vip_patchify(_p(video), _DT[video.dtype], _p(patches), frames, video.shape[-2], video.shape[-1], patch,
                                _stream()), "xp_vip_patchify")


CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)   # dataloader.py:213-214


def vip_patchify_u8(frames_hwc: torch.Tensor, patches: torch.Tensor, patch: int, mean=CLIP_MEAN, std=CLIP_STD):
    """frames_hwc uint8 [..., H, W, 3] -> normalised bf16 patch matrix (the reference's /255 + Normalize fused in)."""
    assert frames_hwc.dtype == torch.uint8 and frames_hwc.is_contiguous() and frames_hwc.shape[-1] == 3
    H, W = frames_hwc.shape[-3], frames_hwc.shape[-2]
    n = frames_hwc.numel() // (3 * H * W)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    check(lib().xp_vip_patchify_u8(_p(frames_hwc), _p(patches), n, H, W, patch, m3, s3, _stream()), "xp_vip_patchify_u8")


def vip_embed_tables(pos, temporal, cls, added, table, x, B, T, L, M, C_, temporal_size):
    check(lib().xp_vip_embed_tables(_p(pos), _p(temporal), _p(cls), _p(added), _p(table), _p(x), B, T, L, M, C_,
                                    temporal_size, _stream()), "xp_vip_embed_tables")


def vip_embed_bwd(d_patch, d_global, d_pos, d_temporal, d_cls, d_added, B, T, L, M, C_, temporal_size):
    check(lib().xp_vip_embed_bwd(_p(d_patch), _p(d_global), _p(d_pos), _p(d_temporal), _p(d_cls), _p(d_added), B, T, L, M,
                                 C_, temporal_size, _stream()), "xp_vip_embed_bwd")


def text_embed_fwd(ids, tok, pos, x, Lt, err_flag):
    rows = ids.numel()
    check(lib().xp_text_embed_fwd(_p(ids), _p(tok), _p(pos), _p(x), rows, Lt, tok.shape[1], tok.shape[0], _p(err_flag),
                                  _stream()), "xp_text_embed_fwd")


def text_embed_bwd(ids, dx, d_tok, d_pos, Lt, C_, vocab):
    check(lib().xp_text_embed_bwd(_p(ids), _p(dx), _p(d_tok), _p(d_pos), ids.numel(), Lt, C_, vocab, _stream()),
          "xp_text_embed_bwd")


def eos_offsets(ids, offsets, index, C_):
    B, Lt = ids.shape
    check(lib().xp_eos_offsets(_p(ids), _p(offsets), _p(index), B, Lt, C_, _stream()), "xp_eos_offsets")


# -------------------------------------------------------------------------------------- attention
def vip_attention_workspace(B, H, T, M, device) -> torch.Tensor:
    n = int(lib().xp_vip_attention_workspace_bytes(B, H, T, M))
    return torch.empty(n // 4, dtype=f32, device=device)


def vip_attention_fwd(qkv, out, lse, ws, B, H, T, L, M, C_):
    check(lib().xp_vip_attention_fwd(_p(qkv), _p(out), _p(lse), _p(ws), B, H, T, L, M, C_, _stream()),
          "xp_vip_attention_fwd")


def vip_attention_fwd_tc(qkv, out, lse, ws, B, H, T, L, M, C_):
    check(lib().xp_vip_attention_fwd_tc(_p(qkv), _p(out), _p(lse), _p(ws), B, H, T, L, M, C_, _stream()),
          "xp_vip_attention_fwd_tc")


def vip_attention_bwd(qkv, out, dout, lse, dqkv, ws, B, H, T, L, M, C_, q_scale):
    check(lib().xp_vip_attention_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _p(ws), B, H, T, L, M, C_, q_scale,
                                     _stream()), "xp_vip_attention_bwd")


def vip_attention_bwd_tc(qkv, out, dout, lse, dqkv, ws, delta, B, H, T, L, M, C_, q_scale):
    check(lib().xp_vip_attention_bwd_tc(_p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _p(ws), _p(delta), B, H, T, L, M,
                                        C_, q_scale, _stream()), "xp_vip_attention_bwd_tc")


def text_attention_fwd(qkv, mask, out, probs, B, H, Lt, C_):
    check(lib().xp_text_attention_fwd(_p(qkv), _p(mask), _p(out), _p(probs), B, H, Lt, C_, _stream()),
          "xp_text_attention_fwd")


def text_attention_bwd(qkv, dout, probs, dqkv, B, H, Lt, C_, q_scale):
    check(lib().xp_text_attention_bwd(_p(qkv), _p(dout), _p(probs), _p(dqkv), B, H, Lt, C_, q_scale, _stream()),
          "xp_text_attention_bwd")


# -------------------------------------------------------------------------------------------- NCE
def nce_split(x, x3, hi, pattern: int):
    rows, d = x.shape
    check(lib().xp_nce_split(_p(x), _p(x3), _p(hi), rows, d, pattern, _stream()), "xp_nce_split")


def nce_softmax_grad(z, logit_scale, lse_r, lse_c, g_scaled, loss, d_logit_scale):
    N, ld = z.shape[0], z.stride(0)
    check(lib().xp_nce_softmax_grad(_p(z), _p(logit_scale), _p(lse_r), _p(lse_c), _p(g_scaled), _p(loss),
                                    _p(d_logit_scale), N, ld, _stream()), "xp_nce_softmax_grad")


def nce_vsc_fc(za, zb, zd, logit_scale, stats, ga, gb, gd, loss, d_logit_scale):
    N, ld = za.shape[0], za.stride(0)
    assert zb.stride(0) == ld and zd.stride(0) == ld and ga.stride(0) == ld and stats.numel() >= 6 * N
    check(lib().xp_nce_vsc_fc(_p(za), _p(zb), _p(zd), _p(logit_scale), _p(stats), _p(ga), _p(gb), _p(gd), _p(loss),
                              _p(d_logit_scale), N, ld, _stream()), "xp_nce_vsc_fc")


# ------------------------------------------------------------- config #4: TimeSformer (HD-VILA)
def seg_desc(n_rows: int, heads: int, ld_qkv: int, ld_out: int, *, n_seq: int, seq_len: int, seg_len: int, inner: int,
             outer_stride: int, inner_stride: int, tok_stride: int) -> XpSegAttn:
    d = XpSegAttn()
    d.n_rows, d.ld_qkv, d.ld_out = n_rows, ld_qkv, ld_out
    d.outer_stride, d.inner_stride, d.tok_stride = outer_stride, inner_stride, tok_stride
    d.heads, d.n_seq, d.seq_len, d.seg_len, d.inner, d.reserved = heads, n_seq, seq_len, seg_len, inner, 0
    return d


def window_desc(n_rows: int, heads: int, head_dim: int, ld_qkv: int, ld_out: int, row_index: torch.Tensor,
                bias: Optional[torch.Tensor], ds_out: Optional[torch.Tensor] = None) -> XpSegAttn:
    """Window attention of LF-VILA's Swin-3D (video_encoder.py:135-164,214-243): row_index int32 [n_windows, L] gives the token
    row of every window position; bias fp32 [nW, heads, L, L] is the relative-position bias (+ shift mask per window type).
    The descriptor keeps references to the tensors (their memory must outlive the launches)."""
    n_win, L = row_index.shape
    assert row_index.dtype == torch.int32 and row_index.is_contiguous()
    d = seg_desc(n_rows, heads, ld_qkv, ld_out, n_seq=n_win, seq_len=L, seg_len=L, inner=1, outer_stride=0, inner_stride=0,
                 tok_stride=1)
    d.row_index = _p(row_index)
    d.head_dim = head_dim
    if bias is not None:
        assert bias.dtype == f32 and bias.is_contiguous() and bias.shape[1:] == (heads, L, L)
        d.bias, d.bias_windows = _p(bias), bias.shape[0]
    if ds_out is not None:
        assert ds_out.dtype == bf16 and ds_out.is_contiguous() and ds_out.shape == (n_win, heads, L, L)
        d.ds_out = _p(ds_out)
    d._keep = (row_index, bias, ds_out)
    return d


def temporal_desc(n_rows: int, T: int, heads: int, ld_qkv: int, ld_out: int) -> XpSegAttn:
    """'(b h w) t m' groups of timesformer.py:210: T consecutive rows each; 64 // T groups share one CTA tile."""
    G = max(1, 64 // T)
    return seg_desc(n_rows, heads, ld_qkv, ld_out, n_seq=(n_rows + G * T - 1) // (G * T), seq_len=G * T, seg_len=T,
                    inner=1, outer_stride=G * T, inner_stride=0, tok_stride=1)


def spatial_desc(B: int, T: int, HW: int, heads: int, ld_qkv: int, ld_out: int) -> XpSegAttn:
    """'(b t) (h w) m' groups of timesformer.py:217: H*W tokens, T rows apart."""
    return seg_desc(B * HW * T, heads, ld_qkv, ld_out, n_seq=B * T, seq_len=HW, seg_len=HW, inner=T,
                    outer_stride=HW * T, inner_stride=1, tok_stride=T)


def seg_attention_fwd(qkv, out, lse, desc: XpSegAttn):
    check(lib().xp_seg_attention_fwd(_p(qkv), _p(out), _p(lse), C.byref(desc), _stream()), "xp_seg_attention_fwd")


def seg_attention_bwd(qkv, out, dout, lse, delta, dqkv, desc: XpSegAttn, q_scale: float):
    check(lib().xp_seg_attention_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(delta), _p(dqkv), C.byref(desc), q_scale,
                                     _stream()), "xp_seg_attention_bwd")


def tsf_embed_fwd(x, pos, time, tokens, B, T, C_, HW):
    check(lib().xp_tsf_embed_fwd(_p(x), _DT[x.dtype], _p(pos), _p(time), _p(tokens), B, T, C_, HW, _stream()),
          "xp_tsf_embed_fwd")


def rowscale(x, scale, out, residual=None):
    """out = (residual or 0) + scale[:, None] * x  (DropPath on a residual branch); out may alias x."""
    rows, C_ = x.shape
    assert scale.dtype == f32 and scale.numel() == rows and x.is_contiguous() and out.is_contiguous()
    check(lib().xp_rowscale_bf16(_p(x), _p(scale), _p(residual), _p(out), rows, C_, _stream()), "xp_rowscale_bf16")


def layernorm_any_fwd(x, y, gamma, beta, mean, rstd, rows: int, C_: int, eps: float):
    """LayerNorm over contiguous [rows, C] (the wide kernel above 1024 columns)."""
    if C_ <= 1024:
        m = rowmap(C_)
        layernorm_fwd(x, m, y, m, gamma, beta, mean, rstd, rows, C_, eps)
    else:
        check(lib().xp_layernorm_wide_fwd(_p(x), _p(y), _p(gamma), _p(beta), _p(mean), _p(rstd), rows, C_, eps, _stream()),
              "xp_layernorm_wide_fwd")


def layernorm_any_bwd(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows: int, C_: int):
    if C_ <= 1024:
        m = rowmap(C_)
        layernorm_bwd(dy, m, x, m, gamma, mean, rstd, dres, m if dres is not None else None, dx, m, dgamma, dbeta, rows, C_)
    else:
        assert dres is None
        check(lib().xp_layernorm_wide_bwd(_p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dgamma), _p(dbeta), rows,
                                          C_, _stream()), "xp_layernorm_wide_bwd")


def gather_rows(src, index, out, C_: int):
    """out.view(-1, C)[i] = src[index[i]] (zeros for index < 0); index int32."""
    assert index.dtype == torch.int32 and index.is_contiguous() and out.numel() == index.numel() * C_
    check(lib().xp_gather_rows_bf16(_p(src), _p(index), _p(out), index.numel(), C_, _stream()), "xp_gather_rows_bf16")


def scatter_rows(inp, index, dst, C_: int):
    """dst[index[i]] = inp.view(-1, C)[i] for index >= 0."""
    assert index.dtype == torch.int32 and index.is_contiguous() and inp.numel() == index.numel() * C_
    check(lib().xp_scatter_rows_bf16(_p(inp), _p(index), _p(dst), index.numel(), C_, _stream()), "xp_scatter_rows_bf16")


def tsf_untokenize(tokens, x, B, T, C_, HW):
    check(lib().xp_tsf_untokenize(_p(tokens), _p(x), _DT[x.dtype], B, T, C_, HW, _stream()), "xp_tsf_untokenize")
