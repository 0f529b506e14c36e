"""B200-native CLIP-ViP dual encoder (video tower with video-proxy tokens + CLIP text tower).

Drop-in for the reference's `CLIPModel` on the VidCLIP path (CLIP-ViP/src/modeling/CLIP_ViP.py): the
module tree below exists only to hold parameters under the reference's exact `state_dict()` names
(`vision_model.pre_layrnorm` spelling included, q/k/v kept as separate parameters), so
released checkpoints, `named_parameters()`-driven weight-decay groups and `load_state_dict_with_mismatch`
work unchanged.  None of these nn.Modules' own forward() is ever called: forward and backward run in one
`torch.autograd.Function` that drives the hand-written sm_100a kernels through the C ABI (xpretrain_b200.ops).
There is no eager / CPU fallback.

Reference call stack replaced:
  CLIPModel.forward            CLIP_ViP.py:1089-1172
  CLIPVisionTransformer.forward :861-903, CLIPVisionViPEmbeddings.forward :168-197
  CLIPTextTransformer.forward  :726-786, CLIPTextEmbeddings.forward :210-227
  CLIPEncoderLayer.forward     :445-460, CLIPAttention.forward2 :332-381 / .forward :266-330, CLIPMLP.forward :392-396
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib, ops
from ._weights import WeightMirror

bf16, f32 = torch.bfloat16, torch.float32


@dataclass
class TowerConfig:
    hidden_size: int
    num_attention_heads: int
    num_hidden_layers: int
    intermediate_size: int


@dataclass
class ClipVipConfig:
    """openai/clip-vit-base-patch16 hyper-parameters + `vision_additional_config` (VidCLIP.py:11-27)."""

    vision: TowerConfig = field(default_factory=lambda: TowerConfig(768, 12, 12, 3072))
    text: TowerConfig = field(default_factory=lambda: TowerConfig(512, 8, 12, 2048))
    image_size: int = 224
    patch_size: int = 16
    projection_dim: int = 512
    vocab_size: int = 49408
    max_position_embeddings: int = 77
    layer_norm_eps: float = 1e-5
    residual_fp32: bool = True          # keep the residual stream out of bf16 (as the reference does under bf16 autocast)
    residual_dtype: str = "fp32"        # its storage type: "fp32", or "fp16" (apex-O2-like: bf16's HBM cost, 8x finer rounding)
    temporal_size: int = 12
    if_use_temporal_embed: int = 1
    add_cls_num: int = 3
    logit_scale_init_value: float = 4.60

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2


# ----------------------------------------------------------------------- parameter containers
class _Attention(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.k_proj = nn.Linear(width, width)
        self.v_proj = nn.Linear(width, width)
        self.q_proj = nn.Linear(width, width)
        self.out_proj = nn.Linear(width, width)


class _MLP(nn.Module):
    def __init__(self, width, inner):
        super().__init__()
        self.fc1 = nn.Linear(width, inner)
        self.fc2 = nn.Linear(inner, width)


class _EncoderLayer(nn.Module):
    def __init__(self, tc: TowerConfig, eps):
        super().__init__()
        self.self_attn = _Attention(tc.hidden_size)
        self.layer_norm1 = nn.LayerNorm(tc.hidden_size, eps=eps)
        self.mlp = _MLP(tc.hidden_size, tc.intermediate_size)
        self.layer_norm2 = nn.LayerNorm(tc.hidden_size, eps=eps)


class _Encoder(nn.Module):
    def __init__(self, tc: TowerConfig, eps):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(tc, eps) for _ in range(tc.num_hidden_layers)])


class _VisionViPEmbeddings(nn.Module):
    def __init__(self, cfg: ClipVipConfig):
        super().__init__()
        w = cfg.vision.hidden_size
        self.added_cls = nn.Parameter(torch.randn(cfg.add_cls_num, w))
        self.class_embedding = nn.Parameter(torch.randn(w))
        self.patch_embedding = nn.Conv2d(3, w, kernel_size=cfg.patch_size, stride=cfg.patch_size, bias=False)
        self.position_embedding = nn.Embedding(cfg.num_patches + 1, w)
        self.register_buffer("position_ids", torch.arange(cfg.num_patches + 1).expand((1, -1)))
        if cfg.if_use_temporal_embed:
            self.temporal_embedding = nn.Parameter(torch.zeros(1, cfg.temporal_size, w))


class _VisionTransformer(nn.Module):
    def __init__(self, cfg: ClipVipConfig):
        super().__init__()
        self.embeddings = _VisionViPEmbeddings(cfg)
        self.pre_layrnorm = nn.LayerNorm(cfg.vision.hidden_size, eps=cfg.layer_norm_eps)  # sic (reference spelling)
        self.encoder = _Encoder(cfg.vision, cfg.layer_norm_eps)
        self.post_layernorm = nn.LayerNorm(cfg.vision.hidden_size, eps=cfg.layer_norm_eps)


class _TextEmbeddings(nn.Module):
    def __init__(self, cfg: ClipVipConfig):
        super().__init__()
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.text.hidden_size)
        self.position_embedding = nn.Embedding(cfg.max_position_embeddings, cfg.text.hidden_size)
        self.register_buffer("position_ids", torch.arange(cfg.max_position_embeddings).expand((1, -1)))


class _TextTransformer(nn.Module):
    def __init__(self, cfg: ClipVipConfig):
        super().__init__()
        self.embeddings = _TextEmbeddings(cfg)
        self.encoder = _Encoder(cfg.text, cfg.layer_norm_eps)
        self.final_layer_norm = nn.LayerNorm(cfg.text.hidden_size, eps=cfg.layer_norm_eps)


def _tower_param_list(tower_prefix: str, n_layers: int) -> List[str]:
    names = []
    for i in range(n_layers):
        p = f"{tower_prefix}.encoder.layers.{i}."
        for lin in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.out_proj", "mlp.fc1", "mlp.fc2"):
            names += [p + lin + ".weight", p + lin + ".bias"]
        for ln in ("layer_norm1", "layer_norm2"):
            names += [p + ln + ".weight", p + ln + ".bias"]
    return names


class CLIPModel(nn.Module):
    """Same constructor argument style, attribute names and output keys as the reference CLIPModel."""

    def __init__(self, config: ClipVipConfig):
        super().__init__()
        self.config = config
        self.vision_model = _VisionTransformer(config)
        self.text_model = _TextTransformer(config)
        self.visual_projection = nn.Linear(config.vision.hidden_size, config.projection_dim, bias=False)
        self.text_projection = nn.Linear(config.text.hidden_size, config.projection_dim, bias=False)
        self.logit_scale = nn.Parameter(torch.ones([]) * config.logit_scale_init_value)
        self._init_weights()
        self._packs: Dict[str, "_WeightPack"] = {}
        # fixed parameter order handed to the autograd.Function (position_ids buffers excluded)
        self._pnames = [n for n, _ in self.named_parameters() if n != "logit_scale"]

    @torch.no_grad()
    def _init_weights(self):
        """CLIPPreTrainedModel._init_weights, CLIP_ViP.py:481-522 (initializer_factor 1, initializer_range 0.02)."""
        cfg = self.config
        te = self.text_model.embeddings
        te.token_embedding.weight.normal_(0.0, 0.02)
        te.position_embedding.weight.normal_(0.0, 0.02)
        ve = self.vision_model.embeddings
        ve.class_embedding.normal_(0.0, cfg.vision.hidden_size ** -0.5)
        ve.patch_embedding.weight.normal_(0.0, 0.02)
        ve.position_embedding.weight.normal_(0.0, 0.02)
        for tower, tc in ((self.vision_model, cfg.vision), (self.text_model, cfg.text)):
            in_std = tc.hidden_size ** -0.5 * (2 * tc.num_hidden_layers) ** -0.5
            for layer in tower.encoder.layers:
                for lin in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj):
                    lin.weight.normal_(0.0, in_std)
                layer.self_attn.out_proj.weight.normal_(0.0, tc.hidden_size ** -0.5)
                layer.mlp.fc1.weight.normal_(0.0, (2 * tc.hidden_size) ** -0.5)
                layer.mlp.fc2.weight.normal_(0.0, in_std)
        self.text_projection.weight.normal_(0.0, cfg.text.hidden_size ** -0.5)
        self.visual_projection.weight.normal_(0.0, cfg.vision.hidden_size ** -0.5)
        for m in self.modules():
            if isinstance(m, nn.LayerNorm):
                m.bias.zero_()
                m.weight.fill_(1.0)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.zero_()

    # ------------------------------------------------------------------------------ public API
    def forward(self, input_ids=None, pixel_values=None, attention_mask=None, return_loss=False, **_unused):
        """CLIPModel.forward, CLIP_ViP.py:1089-1172 (return_loss is accepted and ignored, as VidCLIP passes False)."""
        image_embeds, text_embeds = _run(self, pixel_values, input_ids, attention_mask)
        return {"image_embeds": image_embeds, "text_embeds": text_embeds}

    def get_image_features(self, pixel_values=None, if_norm=None, **_unused):
        """CLIP_ViP.py:1043-1085: projected (and, if if_norm, L2-normalised) video features."""
        image_embeds, _ = _run(self, pixel_values, None, None, normalize=bool(if_norm))
        return image_embeds

    def get_text_features(self, input_ids=None, attention_mask=None, if_norm=None, **_unused):
        """CLIP_ViP.py:992-1041."""
        _, text_embeds = _run(self, None, input_ids, attention_mask, normalize=bool(if_norm))
        return text_embeds


# -------------------------------------------------------------------- bf16 compute copies
class _WeightPack:
    """bf16 compute copies of one tower's GEMM weights with q/k/v fused to [3C, C] (+ the fused fp32 qkv bias)."""

    def __init__(self, layers, device, heads):
        L = len(layers)
        C_ = layers[0].self_attn.q_proj.weight.shape[0]
        I = layers[0].mlp.fc1.weight.shape[0]
        self.C, self.I, self.L = C_, I, L
        self.q_scale = float(C_ // heads) ** -0.5          # CLIPAttention.scale = head_dim ** -0.5 (CLIP_ViP.py:245)
        self.wqkv = torch.empty(L, 3 * C_, C_, dtype=bf16, device=device)
        self.wo = torch.empty(L, C_, C_, dtype=bf16, device=device)
        self.w1 = torch.empty(L, I, C_, dtype=bf16, device=device)
        self.w2 = torch.empty(L, C_, I, dtype=bf16, device=device)
        self.bqkv = torch.empty(L, 3 * C_, dtype=f32, device=device)

    def items(self, layers):
        C_ = self.C
        out = []
        for i, layer in enumerate(layers):
            a = layer.self_attn
            for j, lin in enumerate((a.q_proj, a.k_proj, a.v_proj)):
                out.append((lin.weight, self.wqkv[i, j * C_:(j + 1) * C_]))
                out.append((lin.bias, self.bqkv[i, j * C_:(j + 1) * C_]))
            out += [(a.out_proj.weight, self.wo[i]), (layer.mlp.fc1.weight, self.w1[i]), (layer.mlp.fc2.weight, self.w2[i])]
        return out


def _refresh_weights(model: CLIPModel) -> None:
    """Re-cast EVERY bf16 compute copy from the fp32 masters with one launch (modeling/_weights.py explains why this is
    unconditional: in-place `p.data` writes of the reference's own optimizer must reach the next forward)."""
    dev = model.logit_scale.device
    pk = model._packs
    if pk.get("device") != dev:
        pk.clear()
        pk["device"] = dev
        pk["vision"] = _WeightPack(model.vision_model.encoder.layers, dev, model.config.vision.num_attention_heads)
        pk["text"] = _WeightPack(model.text_model.encoder.layers, dev, model.config.text.num_attention_heads)
        for key, p in (("small:patch", model.vision_model.embeddings.patch_embedding.weight),
                       ("small:vproj", model.visual_projection.weight), ("small:tproj", model.text_projection.weight)):
            pk[key] = torch.empty(p.shape, dtype=bf16, device=dev)
        pk["mirror"] = WeightMirror()
        items = pk["vision"].items(model.vision_model.encoder.layers) + pk["text"].items(model.text_model.encoder.layers)
        items += [(model.vision_model.embeddings.patch_embedding.weight, pk["small:patch"]),
                  (model.visual_projection.weight, pk["small:vproj"]), (model.text_projection.weight, pk["small:tproj"])]
        pk["items"] = items           # (Parameter, destination view) pairs: Parameter objects are stable, their .data may move
    pk["mirror"].refresh(pk["items"])


def _pack(model: CLIPModel, which: str) -> _WeightPack:
    return model._packs[which]


def _small_bf16(model: CLIPModel, name: str, p: torch.Tensor) -> torch.Tensor:
    return model._packs["small:" + name]


# --------------------------------------------------------------------------- encoder layers
def _residual_fp32(model) -> bool:
    """The residual stream is kept OUT of bf16 (as under the reference's bf16 autocast, where only the Linear / matmul inputs are
    rounded): the block outputs stay bf16 branch tensors and the add happens in fp32 inside the next LayerNorm kernel.
    `model.config.residual_fp32 = False` (or XP_RESIDUAL_BF16=1) selects the round-1 path: bf16 stream, add in the GEMM epilogue."""
    import os
    return bool(getattr(model.config, "residual_fp32", True)) and os.environ.get("XP_RESIDUAL_BF16") != "1"


def _stream_dtype(model) -> torch.dtype:
    """Storage type of the residual stream between the fused add + LayerNorm kernels: fp32 (default), or fp16
    (`config.residual_dtype = "fp16"` / XP_RESIDUAL_DTYPE=fp16): 11 mantissa bits instead of bf16's 8 at bf16's HBM cost — the
    precision the reference itself trains in under apex O2 (run_pretrain.py:234-236); values saturate at +-65504."""
    import os
    name = os.environ.get("XP_RESIDUAL_DTYPE") or getattr(model.config, "residual_dtype", "fp32")
    return torch.float16 if str(name) in ("fp16", "float16", "half") else f32


def _layer_fwd(x, pend, layer, pk: _WeightPack, i: int, eps: float, attn_fwd, rows: int, save: bool, stream_dt):
    """One pre-LN residual block (CLIP_ViP.py:445-460).  x: residual stream [rows, C] in `stream_dt` (fp32 / fp16), or bf16 when
    stream_dt is None (round-1 path); pend: the previous block's bf16 branch output that still has to be added to it.
    Returns (x_out, pend_out, saved)."""
    fp32res = stream_dt is not None
    C_, I = pk.C, pk.I
    dev = x.device
    plain = ops.rowmap(C_)
    mean1 = torch.empty(rows, dtype=f32, device=dev); rstd1 = torch.empty_like(mean1)
    h = torch.empty(rows, C_, dtype=bf16, device=dev)
    ln1, ln2 = layer.layer_norm1, layer.layer_norm2
    if pend is not None:        # x <- x + pend in fp32, fused into layer_norm1
        xs = torch.empty(rows, C_, dtype=stream_dt, device=dev)
        ops.layernorm_fwd(x, plain, h, plain, ln1.weight, ln1.bias, mean1, rstd1, rows, C_, eps, add=pend, addmap=plain,
                          sum_out=xs, summap=plain)
        x = xs
    else:
        ops.layernorm_fwd(x, plain, h, plain, ln1.weight, ln1.bias, mean1, rstd1, rows, C_, eps)
    qkv = torch.empty(rows, 3 * C_, dtype=bf16, device=dev)
    # q = (h Wq^T + bq) * head_dim**-0.5 : the scale multiplies the bias too (CLIP_ViP.py:341 / :269)
    ops.linear_fwd(h, pk.wqkv[i], pk.bqkv[i], qkv, scale_cols=C_, col_scale=pk.q_scale)
    a = torch.empty(rows, C_, dtype=bf16, device=dev)
    att_saved = attn_fwd(qkv, a)
    mean2 = torch.empty(rows, dtype=f32, device=dev); rstd2 = torch.empty_like(mean2)
    h2 = torch.empty(rows, C_, dtype=bf16, device=dev)
    if fp32res:
        y1 = torch.empty(rows, C_, dtype=bf16, device=dev)
        ops.linear_fwd(a, pk.wo[i], layer.self_attn.out_proj.bias, y1)                  # branch only: the add is in layer_norm2
        x1 = torch.empty(rows, C_, dtype=stream_dt, device=dev)
        ops.layernorm_fwd(x, plain, h2, plain, ln2.weight, ln2.bias, mean2, rstd2, rows, C_, eps, add=y1, addmap=plain,
                          sum_out=x1, summap=plain)
        del y1
    else:
        x1 = torch.empty(rows, C_, dtype=bf16, device=dev)
        ops.linear_fwd(a, pk.wo[i], layer.self_attn.out_proj.bias, x1, residual=x, ldr=C_)
        ops.layernorm_fwd(x1, plain, h2, plain, ln2.weight, ln2.bias, mean2, rstd2, rows, C_, eps)
    pre = torch.empty(rows, I, dtype=bf16, device=dev) if save else None
    f1 = torch.empty(rows, I, dtype=bf16, device=dev)
    ops.linear_fwd(h2, pk.w1[i], layer.mlp.fc1.bias, f1, act=_lib.ACT_QUICK_GELU, aux=pre, ld_aux=I)
    out = torch.empty(rows, C_, dtype=bf16, device=dev)
    saved = (x, mean1, rstd1, h, qkv, att_saved, a, x1, mean2, rstd2, h2, pre, f1) if save else None
    if fp32res:
        ops.linear_fwd(f1, pk.w2[i], layer.mlp.fc2.bias, out)                           # branch; added by the next LayerNorm
        return x1, out, saved
    ops.linear_fwd(f1, pk.w2[i], layer.mlp.fc2.bias, out, residual=x1, ldr=C_)
    return out, None, saved


def _pooled_ln(x, pend, rmap, ln, B: int, C_: int, eps: float):
    """LayerNorm of B selected rows (CLS / EOS, picked by `rmap`) of the final hidden state x (+ pend) -> (pooled bf16 [B, C],
    mean, rstd, (saved LayerNorm input, its row map))."""
    dev = x.device
    plain = ops.rowmap(C_)
    pooled = torch.empty(B, C_, dtype=bf16, device=dev)
    mean = torch.empty(B, dtype=f32, device=dev); rstd = torch.empty_like(mean)
    if pend is None:
        ops.layernorm_fwd(x, rmap, pooled, plain, ln.weight, ln.bias, mean, rstd, B, C_, eps)
        return pooled, mean, rstd, (x, rmap)
    rows_in = torch.empty(B, C_, dtype=torch.float16 if x.dtype == torch.float16 else f32, device=dev)
    ops.layernorm_fwd(x, rmap, pooled, plain, ln.weight, ln.bias, mean, rstd, B, C_, eps, add=pend, addmap=rmap, sum_out=rows_in,
                      summap=plain)
    return pooled, mean, rstd, (rows_in, plain)


def _colsum(x: torch.Tensor, out: torch.Tensor, aux) -> None:
    """Bias gradient = column sum of x (HBM-bound).  With an auxiliary stream it runs UNDER the tensor-bound GEMMs that follow
    (its 256-thread CTAs co-reside with the persistent GEMM CTAs); the caller joins the stream before the gradients are used."""
    if aux is None:
        ops.colsum(x, out)
        return
    aux.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(aux):
        ops.colsum(x, out)
    x.record_stream(aux)


def _layer_bwd(dx, saved, layer, pk: _WeightPack, i: int, grads: Dict[str, torch.Tensor], prefix: str, attn_bwd,
               rows: int, aux=None):
    """Backward of one block; dx [rows, C] bf16 is d(loss)/d(block output).  Returns d(block input)."""
    (x, mean1, rstd1, h, qkv, att_saved, a, x1, mean2, rstd2, h2, pre, f1) = saved
    C_, I = pk.C, pk.I
    dev = dx.device
    plain = ops.rowmap(C_)
    g = lambda n: grads[prefix + n]  # noqa: E731
    # ---- x_out = x1 + fc2(quick_gelu(fc1(LN2(x1))))
    ops.linear_wgrad(dx, f1, g("mlp.fc2.weight"))
    dpre = torch.empty(rows, I, dtype=bf16, device=dev)
    ops.linear_dgrad(dx, pk.w2[i], dpre, act=_lib.ACT_DQUICK_GELU, aux=pre, ld_aux=I)
    _colsum(dpre, g("mlp.fc1.bias"), aux)
    ops.linear_wgrad(dpre, h2, g("mlp.fc1.weight"))
    dh2 = torch.empty(rows, C_, dtype=bf16, device=dev)
    ops.linear_dgrad(dpre, pk.w1[i], dh2)
    del dpre
    dx1 = torch.empty(rows, C_, dtype=bf16, device=dev)
    # LN2 backward streams dx as the residual-branch gradient: its column sum (= fc2.bias gradient) comes out of the same pass
    ops.layernorm_bwd(dh2, plain, x1, plain, layer.layer_norm2.weight, mean2, rstd2, dx, plain, dx1, plain,
                      g("layer_norm2.weight"), g("layer_norm2.bias"), rows, C_, dres_colsum=g("mlp.fc2.bias"))
    # ---- x1 = x + out_proj(attn(qkv(LN1(x))))
    ops.linear_wgrad(dx1, a, g("self_attn.out_proj.weight"))
    da = dh2  # reuse
    ops.linear_dgrad(dx1, pk.wo[i], da)
    dqkv = torch.empty(rows, 3 * C_, dtype=bf16, device=dev)
    attn_bwd(qkv, a, da, att_saved, dqkv)
    dwqkv = grads[prefix + "self_attn.qkv.weight"]
    dbqkv = grads[prefix + "self_attn.qkv.bias"]
    _colsum(dqkv, dbqkv, aux)
    ops.linear_wgrad(dqkv, h, dwqkv)
    dh = da
    ops.linear_dgrad(dqkv, pk.wqkv[i], dh)
    del dqkv
    dxin = torch.empty(rows, C_, dtype=bf16, device=dev)
    ops.layernorm_bwd(dh, plain, x, plain, layer.layer_norm1.weight, mean1, rstd1, dx1, plain, dxin, plain,
                      g("layer_norm1.weight"), g("layer_norm1.bias"), rows, C_, dres_colsum=g("self_attn.out_proj.bias"))
    if aux is not None:
        torch.cuda.current_stream().wait_stream(aux)      # this layer's bias gradients are complete
    return dxin


def _alloc_flat(shapes: Dict[str, tuple], grads: Dict[str, torch.Tensor], dev) -> torch.Tensor:
    """One zeroed fp32 buffer holding all gradients of a group as 16-byte aligned views: the data-parallel
    all-reduce of the group is then a single collective on `flat` (no packing copies)."""
    offs, total = {}, 0
    for n, shp in shapes.items():
        offs[n] = total
        total += (int(torch.Size(shp).numel()) + 3) // 4 * 4
    flat = torch.zeros(total, dtype=f32, device=dev)
    for n, shp in shapes.items():
        grads[n] = flat[offs[n]:offs[n] + torch.Size(shp).numel()].view(shp)
    return flat


def _alloc_layer_grads(layer, prefix: str, grads: Dict[str, torch.Tensor], dev) -> torch.Tensor:
    C_ = layer.self_attn.q_proj.weight.shape[0]
    shapes = {prefix + "self_attn.qkv.weight": (3 * C_, C_), prefix + "self_attn.qkv.bias": (3 * C_,)}
    for n, p in layer.named_parameters():
        if ".q_proj." in n or ".k_proj." in n or ".v_proj." in n:
            continue
        shapes[prefix + n] = tuple(p.shape)
    return _alloc_flat(shapes, grads, dev)


def _grads_ready(model, grads: Dict[str, torch.Tensor], key: str) -> None:
    """A gradient group (one encoder layer, or a tower's remaining parameters) is final: hand its flat buffer to
    `model.grad_ready_hook` (e.g. utils.distributed.OverlappedGradAverager, which starts an async NCCL all-reduce
    that overlaps with the rest of the backward pass — the hvd.DistributedOptimizer hooks of run_pretrain.py:226-228)."""
    hook = getattr(model, "grad_ready_hook", None)
    flat = grads.get("__flat__" + key)
    if hook is not None and flat is not None:
        hook(flat)


def _finish_layer_grads(prefix: str, grads: Dict[str, torch.Tensor], C_: int):
    w = grads.pop(prefix + "self_attn.qkv.weight")
    b = grads.pop(prefix + "self_attn.qkv.bias")
    for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
        grads[prefix + f"self_attn.{n}.weight"] = w[j * C_:(j + 1) * C_]
        grads[prefix + f"self_attn.{n}.bias"] = b[j * C_:(j + 1) * C_]


# ------------------------------------------------------------------------------ vision tower
def _vision_fwd(model: CLIPModel, video: torch.Tensor, save: bool):
    cfg = model.config
    vm = model.vision_model
    emb = vm.embeddings
    B, T = video.shape[0], video.shape[1]
    if video.dtype == torch.uint8 and (video.dim() != 5 or video.shape[-1] != 3):
        raise ValueError("uint8 video must be channels-last [B, T, H, W, 3] (decoder layout)")
    C_, L, M = cfg.vision.hidden_size, cfg.num_patches, 1 + cfg.add_cls_num
    H = cfg.vision.num_attention_heads
    S = M + T * L
    rows = B * S
    dev = video.device
    pk = _pack(model, "vision")
    eps = cfg.layer_norm_eps
    Kp = 3 * cfg.patch_size * cfg.patch_size

    patches = torch.empty(B * T * L, Kp, dtype=bf16, device=dev)
    if video.dtype == torch.uint8:      # raw decoder frames: the reference's /255 + Normalize is fused into the patch extraction
        ops.vip_patchify_u8(video.contiguous(), patches, cfg.patch_size, getattr(model, "pixel_mean", ops.CLIP_MEAN),
                            getattr(model, "pixel_std", ops.CLIP_STD))
    else:
        ops.vip_patchify(video.contiguous(), patches, cfg.patch_size)
    table = torch.empty(T * L, C_, dtype=bf16, device=dev)
    x0 = torch.empty(rows, C_, dtype=bf16, device=dev)
    temporal = emb.temporal_embedding if cfg.if_use_temporal_embed else None
    ops.vip_embed_tables(emb.position_embedding.weight, temporal, emb.class_embedding, emb.added_cls, table, x0, B, T, L,
                         M, C_, cfg.temporal_size)
    wp = _small_bf16(model, "patch", emb.patch_embedding.weight).view(C_, Kp)
    # conv-as-GEMM; epilogue adds the periodic [T*L, C] position+temporal table and writes past the M global rows
    ops.gemm(patches, wp, x0, M=B * T * L, N=C_, K=Kp, lda=Kp, ldb=Kp, ldc=C_, residual=table, ldr=C_, r_group=T * L,
             r_group_stride=0, c_group=T * L, c_group_stride=S * C_, c_offset=M * C_)
    plain = ops.rowmap(C_)
    # pre_layrnorm (CLIP_ViP.py:881) in two launches so that its statistics are stored compactly per half
    # (patch rows / global rows), the layout its backward and the embedding backward consume
    pmap = ops.rowmap(C_, group=T * L, group_stride=S * C_)
    gmap = ops.rowmap(C_, group=M, group_stride=S * C_)
    mean0p = torch.empty(B * T * L, dtype=f32, device=dev); rstd0p = torch.empty_like(mean0p)
    mean0g = torch.empty(B * M, dtype=f32, device=dev); rstd0g = torch.empty_like(mean0g)
    stream_dt = _stream_dtype(model) if _residual_fp32(model) else None
    x = torch.empty(rows, C_, dtype=stream_dt if stream_dt is not None else bf16, device=dev)
    ln0 = vm.pre_layrnorm
    ops.layernorm_fwd(x0, pmap, x, pmap, ln0.weight, ln0.bias, mean0p, rstd0p, B * T * L, C_, eps, x_off=M * C_,
                      y_off=M * C_)
    ops.layernorm_fwd(x0, gmap, x, gmap, ln0.weight, ln0.bias, mean0g, rstd0g, B * M, C_, eps)

    ws = ops.vip_attention_workspace(B, H, T, M, dev)

    def attn_fwd(qkv, out):
        lse = torch.empty(B, H, S, dtype=f32, device=dev)
        ops.vip_attention_fwd_tc(qkv, out, lse, ws, B, H, T, L, M, C_)
        return lse

    layer_saved = []
    pend = None
    timer = getattr(model, "block_timer", None)   # bench.py: CUDA events around each ViP block (metric 2 of BASELINE.json)
    for i, layer in enumerate(vm.encoder.layers):
        if timer is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        x, pend, sv = _layer_fwd(x, pend, layer, pk, i, eps, attn_fwd, rows, save, stream_dt)
        if timer is not None:
            e1.record()
            timer.append(("fwd", e0, e1))
        layer_saved.append(sv)
    # pooled = post_layernorm(last_hidden[:, 0])  (CLIP_ViP.py:891-893): CLS rows picked by the row map
    cls_map = ops.rowmap(C_, group=1, group_stride=S * C_)
    pooled, meanp, rstdp, post_in = _pooled_ln(x, pend, cls_map, vm.post_layernorm, B, C_, eps)
    wproj = _small_bf16(model, "vproj", model.visual_projection.weight)
    proj = torch.empty(B, cfg.projection_dim, dtype=f32, device=dev)
    ops.linear_fwd(pooled, wproj, None, proj, out_mode=_lib.OUT_F32)
    saved = None
    if save:
        saved = SimpleNamespace(B=B, T=T, S=S, rows=rows, patches=patches, x0=x0, stats0=(mean0p, rstd0p, mean0g, rstd0g),
                                layers=layer_saved, x_last=(x, pend), post_in=post_in, pooled=pooled, meanp=meanp, rstdp=rstdp,
                                ws=ws)
    return proj, saved


def _vision_bwd(model: CLIPModel, dproj_bf16: torch.Tensor, sv, grads: Dict[str, torch.Tensor]):
    """dproj_bf16 [B, proj] = gradient w.r.t. the un-normalised projection output."""
    cfg = model.config
    vm = model.vision_model
    C_, L, M = cfg.vision.hidden_size, cfg.num_patches, 1 + cfg.add_cls_num
    H = cfg.vision.num_attention_heads
    B, T, S, rows = sv.B, sv.T, sv.S, sv.rows
    dev = dproj_bf16.device
    pk = _pack(model, "vision")
    plain = ops.rowmap(C_)
    wproj = _small_bf16(model, "vproj", model.visual_projection.weight)
    ops.linear_wgrad(dproj_bf16, sv.pooled, grads["visual_projection.weight"])
    dpooled = torch.empty(B, C_, dtype=bf16, device=dev)
    ops.linear_dgrad(dproj_bf16, wproj, dpooled)
    dx = torch.zeros(rows, C_, dtype=bf16, device=dev)  # only the CLS rows receive gradient from the head
    cls_map = ops.rowmap(C_, group=1, group_stride=S * C_)
    ops.layernorm_bwd(dpooled, plain, sv.post_in[0], sv.post_in[1], vm.post_layernorm.weight, sv.meanp, sv.rstdp, None, None,
                      dx, cls_map, grads["vision_model.post_layernorm.weight"], grads["vision_model.post_layernorm.bias"], B, C_)

    delta = torch.empty(B, H, S, dtype=f32, device=dev)       # rowsum(dO * O), scratch of the attention backward

    def attn_bwd(qkv, a, da, lse, dqkv):
        # pipelined tcgen05 / TMEM backward (csrc/vip_attention_tc.cu); the mma.sync kernel of round 1 stays as a cross-check in tests
        ops.vip_attention_bwd_tc(qkv, a, da, lse, dqkv, sv.ws, delta, B, H, T, L, M, C_, pk.q_scale)

    timer = getattr(model, "block_timer", None)
    aux = _aux_stream(model, dev)
    for i in reversed(range(len(vm.encoder.layers))):
        prefix = f"vision_model.encoder.layers.{i}."
        if timer is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        dx = _layer_bwd(dx, sv.layers[i], vm.encoder.layers[i], pk, i, grads, prefix, attn_bwd, rows, aux)
        if timer is not None:
            e1.record()
            timer.append(("bwd", e0, e1))
        sv.layers[i] = None
        _grads_ready(model, grads, prefix)
    # pre_layrnorm backward, written as two compact halves: patch rows [B, T*L, C] and global rows [B, M, C]
    d_patch = torch.empty(B * T * L, C_, dtype=bf16, device=dev)
    d_glob = torch.empty(B * M, C_, dtype=bf16, device=dev)
    pmap = ops.rowmap(C_, group=T * L, group_stride=S * C_)
    gmap = ops.rowmap(C_, group=M, group_stride=S * C_)
    gw, gb = grads["vision_model.pre_layrnorm.weight"], grads["vision_model.pre_layrnorm.bias"]
    mean0p, rstd0p, mean0g, rstd0g = sv.stats0
    ops.layernorm_bwd(dx, pmap, sv.x0, pmap, vm.pre_layrnorm.weight, mean0p, rstd0p, None, None, d_patch, plain, gw, gb,
                      B * T * L, C_, dy_off=M * C_, x_off=M * C_)
    ops.layernorm_bwd(dx, gmap, sv.x0, gmap, vm.pre_layrnorm.weight, mean0g, rstd0g, None, None, d_glob, plain, gw, gb,
                      B * M, C_)
    emb = "vision_model.embeddings."
    ops.vip_embed_bwd(d_patch, d_glob, grads[emb + "position_embedding.weight"],
                      grads.get(emb + "temporal_embedding"), grads[emb + "class_embedding"], grads[emb + "added_cls"],
                      B, T, L, M, C_, cfg.temporal_size)
    Kp = 3 * cfg.patch_size * cfg.patch_size
    ops.linear_wgrad(d_patch, sv.patches, grads[emb + "patch_embedding.weight"].view(C_, Kp))


# -------------------------------------------------------------------------------- text tower
def _text_fwd(model: CLIPModel, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], save: bool):
    cfg = model.config
    tm = model.text_model
    B, Lt = input_ids.shape
    C_, H = cfg.text.hidden_size, cfg.text.num_attention_heads
    rows = B * Lt
    dev = input_ids.device
    pk = _pack(model, "text")
    eps = cfg.layer_norm_eps
    if Lt > cfg.max_position_embeddings:      # the reference fails in position_ids[:, :seq_length] + embedding add (CLIP_ViP.py:217-225)
        raise ValueError(f"text length {Lt} exceeds max_position_embeddings {cfg.max_position_embeddings}")
    ids = input_ids.contiguous().to(torch.int64)
    mask = attention_mask.contiguous().to(torch.int64) if attention_mask is not None else None
    x = torch.empty(rows, C_, dtype=bf16, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.text_embed_fwd(ids, tm.embeddings.token_embedding.weight, tm.embeddings.position_embedding.weight, x, Lt, err)
    if getattr(model, "validate_ids", False) and int(err.item()) != 0:   # opt-in: costs a device sync per forward
        raise IndexError("text_input_ids contains a token id outside [0, vocab_size) (nn.Embedding would raise, CLIP_ViP.py:222)")

    def attn_fwd(qkv, out):
        probs = torch.empty(B, H, Lt, Lt, dtype=f32, device=dev)
        ops.text_attention_fwd(qkv, mask, out, probs, B, H, Lt, C_)
        return probs

    layer_saved = []
    pend = None
    stream_dt = _stream_dtype(model) if _residual_fp32(model) else None
    if stream_dt is not None:
        x = x.to(stream_dt)          # [B*Lt, 512]: the token + position embeddings enter the stream in its storage type
    for i, layer in enumerate(tm.encoder.layers):
        x, pend, sv = _layer_fwd(x, pend, layer, pk, i, eps, attn_fwd, rows, save, stream_dt)
        layer_saved.append(sv)
    # final_layer_norm is per-row, so it is applied to the pooled EOS row only (first argmax of the ids, :776)
    eos = torch.empty(B, dtype=torch.int64, device=dev)
    ops.eos_offsets(ids, eos, None, C_)
    plain = ops.rowmap(C_)
    emap = ops.rowmap(C_, offsets=eos)
    pooled, meanp, rstdp, post_in = _pooled_ln(x, pend, emap, tm.final_layer_norm, B, C_, eps)
    wproj = _small_bf16(model, "tproj", model.text_projection.weight)
    proj = torch.empty(B, cfg.projection_dim, dtype=f32, device=dev)
    ops.linear_fwd(pooled, wproj, None, proj, out_mode=_lib.OUT_F32)
    saved = None
    if save:
        saved = SimpleNamespace(B=B, Lt=Lt, rows=rows, ids=ids, layers=layer_saved, x_last=(x, pend), post_in=post_in, eos=eos,
                                pooled=pooled, meanp=meanp, rstdp=rstdp, err=err)
    return proj, saved


def _text_bwd(model: CLIPModel, dproj_bf16: torch.Tensor, sv, grads: Dict[str, torch.Tensor]):
    cfg = model.config
    tm = model.text_model
    C_, H = cfg.text.hidden_size, cfg.text.num_attention_heads
    B, Lt, rows = sv.B, sv.Lt, sv.rows
    dev = dproj_bf16.device
    pk = _pack(model, "text")
    plain = ops.rowmap(C_)
    wproj = _small_bf16(model, "tproj", model.text_projection.weight)
    ops.linear_wgrad(dproj_bf16, sv.pooled, grads["text_projection.weight"])
    dpooled = torch.empty(B, C_, dtype=bf16, device=dev)
    ops.linear_dgrad(dproj_bf16, wproj, dpooled)
    dx = torch.zeros(rows, C_, dtype=bf16, device=dev)
    emap = ops.rowmap(C_, offsets=sv.eos)
    ops.layernorm_bwd(dpooled, plain, sv.post_in[0], sv.post_in[1], tm.final_layer_norm.weight, sv.meanp, sv.rstdp, None, None,
                      dx, emap, grads["text_model.final_layer_norm.weight"], grads["text_model.final_layer_norm.bias"], B, C_)

    def attn_bwd(qkv, a, da, probs, dqkv):
        ops.text_attention_bwd(qkv, da, probs, dqkv, B, H, Lt, C_, pk.q_scale)

    for i in reversed(range(len(tm.encoder.layers))):
        prefix = f"text_model.encoder.layers.{i}."
        dx = _layer_bwd(dx, sv.layers[i], tm.encoder.layers[i], pk, i, grads, prefix, attn_bwd, rows)
        sv.layers[i] = None
        _grads_ready(model, grads, prefix)
    ops.text_embed_bwd(sv.ids, dx, grads["text_model.embeddings.token_embedding.weight"],
                       grads["text_model.embeddings.position_embedding.weight"], Lt, C_, cfg.vocab_size)


# ------------------------------------------------------------------- the autograd.Function
class _ClipVipFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: CLIPModel, video, input_ids, attention_mask, normalize, grad_mode, *params):
        # grad_mode = torch.is_grad_enabled() of the caller (Function.forward itself always runs under no_grad, and
        # needs_input_grad reflects requires_grad even then): evaluation must not keep the activations alive
        names = model._pnames
        need = {n: r for n, r in zip(names, ctx.needs_input_grad[6:])}
        save = grad_mode and any(need.values())
        ctx.model = model
        _refresh_weights(model)
        ctx.normalize = normalize
        ctx.vis = ctx.txt = None
        dev = params[0].device

        def run_tower(which):
            if which == "vis":
                tower_save = save and any(r for n, r in need.items() if n.startswith(("vision_model.", "visual_projection.")))
                proj, sv = _vision_fwd(model, video, tower_save)
            else:
                # a frozen text tower (VidCLIP.freeze_text_encoder) keeps nothing and runs no backward
                tower_save = save and any(r for n, r in need.items() if n.startswith(("text_model.", "text_projection.")))
                proj, sv = _text_fwd(model, input_ids, attention_mask, tower_save)
            if normalize:
                feat = torch.empty_like(proj)
                inv = torch.empty(proj.shape[0], dtype=f32, device=dev)
                ops.l2norm_fwd(proj, feat, inv)
            else:
                feat, inv = proj, None
            if sv is not None:
                sv.feat, sv.inv = feat, inv
                setattr(ctx, which, sv)
            return feat

        none = torch.empty(0, device=dev)
        side = _side_stream(model, dev) if (video is not None and input_ids is not None) else None
        ctx.side = side
        if side is None:
            vis = run_tower("vis") if video is not None else none
            txt = run_tower("txt") if input_ids is not None else none
        else:
            # The text tower (~300 launches of microsecond kernels) runs on a side stream under the vision
            # tower: its small grids fill the SMs that the persistent vision kernels leave idle in their last wave.
            # Issue order: vision first.  After a host sync (a driver reading loss.item() every step) the GPU would otherwise sit on
            # microsecond text kernels while the host is still enqueueing them; queued behind ~600 vision launches they still
            # start while the vision tower is executing.
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            vis = run_tower("vis")
            with torch.cuda.stream(side):
                txt = run_tower("txt")
            main.wait_stream(side)
            txt.record_stream(main)
        return vis, txt

    @staticmethod
    def backward(ctx, d_vis, d_txt):
        model: CLIPModel = ctx.model
        dev = model.logit_scale.device
        names = model._pnames
        named = dict(model.named_parameters())
        grads: Dict[str, torch.Tensor] = {}
        C_v, C_t = model.config.vision.hidden_size, model.config.text.hidden_size
        jobs = (("vision_model", ctx.vis, d_vis, "visual_projection.weight", _vision_bwd, C_v),
                ("text_model", ctx.txt, d_txt, "text_projection.weight", _text_bwd, C_t))
        main = torch.cuda.current_stream()
        side = ctx.side if (ctx.vis is not None and ctx.txt is not None and d_vis is not None and d_txt is not None) else None
        # data-parallel runs: leave `model.nccl_sm_reserve` SMs to the NCCL kernels of the overlapped gradient all-reduce for
        # the duration of the backward pass only (the forward has no collective in flight and keeps every SM)
        reserve = int(getattr(model, "nccl_sm_reserve", 0) or 0)
        if reserve > 0:
            ops.set_sm_limit(torch.cuda.get_device_properties(dev).multi_processor_count - reserve)
        if side is not None:
            side.wait_stream(main)            # before any vision-backward launch: the text backward only needs d_txt
        for tower, sv, dfeat, proj_name, bwd, C_ in jobs:
            if sv is None or dfeat is None:
                continue
            on_side = side is not None and tower == "text_model"
            if on_side:        # text backward on the side stream, under the vision backward (issued after it, see forward)
                dfeat.record_stream(side)
                with torch.cuda.stream(side):
                    _tower_backward(model, ctx, tower, sv, dfeat, proj_name, bwd, C_, grads, names, named, dev)
                continue
            _tower_backward(model, ctx, tower, sv, dfeat, proj_name, bwd, C_, grads, names, named, dev)
        if side is not None:
            main.wait_stream(side)
            for k, t in grads.items():      # allocated in the side stream's pool, consumed by autograd on the main stream
                if k.startswith("__flat__text_model"):
                    t.record_stream(main)
        hook = getattr(model, "grad_ready_hook", None)
        if hook is not None and hasattr(hook, "finish"):
            hook.finish()      # stream-ordered wait: autograd's accumulation below sees the averaged values
        if reserve > 0:
            ops.set_sm_limit(0)
        ctx.vis = ctx.txt = None
        return (None, None, None, None, None, None) + tuple(grads.get(n) if r else None
                                                             for n, r in zip(names, ctx.needs_input_grad[6:]))


def _tower_backward(model, ctx, tower, sv, dfeat, proj_name, bwd, C_, grads, names, named, dev):
    tw = getattr(model, tower)
    for i, layer in enumerate(tw.encoder.layers):
        pre = f"{tower}.encoder.layers.{i}."
        grads["__flat__" + pre] = _alloc_layer_grads(layer, pre, grads, dev)
    rest = {n: tuple(named[n].shape) for n in names if n.startswith(tower + ".") and ".encoder.layers." not in n}
    rest[proj_name] = tuple(named[proj_name].shape)
    grads["__flat__" + tower] = _alloc_flat(rest, grads, dev)
    dproj = torch.empty(dfeat.shape, dtype=bf16, device=dev)
    dfeat = dfeat.contiguous().to(f32)
    if ctx.normalize:
        ops.l2norm_bwd(dfeat, sv.feat, sv.inv, dproj)
    else:
        dproj.copy_(dfeat)
    bwd(model, dproj, sv, grads)
    _grads_ready(model, grads, tower)
    for i in range(len(tw.encoder.layers)):
        _finish_layer_grads(f"{tower}.encoder.layers.{i}.", grads, C_)


def _side_stream(model: CLIPModel, dev):
    """Side stream for the text tower (None when `model.overlap_text_tower` is False or XP_NO_OVERLAP=1)."""
    import os
    if not getattr(model, "overlap_text_tower", True) or os.environ.get("XP_NO_OVERLAP") == "1":
        return None
    st = model._packs.get("side_stream")
    if st is None or st.device != dev:
        st = model._packs["side_stream"] = torch.cuda.Stream(device=dev)
    return st


def _aux_stream(model: CLIPModel, dev):
    """Stream for the HBM-bound bias column sums of the vision backward (None: XP_NO_OVERLAP=1 / model.overlap_colsum False)."""
    import os
    if not getattr(model, "overlap_colsum", True) or os.environ.get("XP_NO_OVERLAP") == "1":
        return None
    st = model._packs.get("aux_stream")
    if st is None or st.device != dev:
        st = model._packs["aux_stream"] = torch.cuda.Stream(device=dev)
    return st


def _run(model: CLIPModel, video, input_ids, attention_mask, normalize: bool = True):
    if not model.logit_scale.is_cuda:
        raise _lib.XpError("xpretrain_b200.CLIPModel must live on a CUDA (B200) device: there is no CPU path")
    params = [p for n, p in model.named_parameters() if n != "logit_scale"]
    vis, txt = _ClipVipFunction.apply(model, video, input_ids, attention_mask, normalize, torch.is_grad_enabled(), *params)
    return (vis if video is not None else None), (txt if input_ids is not None else None)
