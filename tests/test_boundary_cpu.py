"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side logic that needs no GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from xpretrain_b200 import _lib
    header = open(os.path.join(ROOT, "include", "xpretrain_b200.h")).read()
    declared = set(re.findall(r"\b(xp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/xpretrain_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert handle.xp_version() == 1


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of computing on the host."""
    from types import SimpleNamespace
    from xpretrain_b200 import _lib
    from xpretrain_b200.modeling import VidCLIP
    from xpretrain_b200.modeling.clip_vip import ClipVipConfig, TowerConfig
    add = SimpleNamespace(type="ViP", temporal_size=12, if_use_temporal_embed=1, logit_scale_init_value=4.6, add_cls_num=3)
    mc = ClipVipConfig(vision=TowerConfig(768, 12, 1, 3072), text=TowerConfig(512, 8, 1, 2048))
    model = VidCLIP(SimpleNamespace(clip_config=mc, clip_weights="", clip_vision_additional_config=add))
    with pytest.raises(_lib.XpError):
        model(video=torch.zeros(1, 1, 3, 224, 224), text_input_ids=torch.zeros(1, 4, dtype=torch.long),
              text_input_mask=torch.ones(1, 4, dtype=torch.long))


def test_product_path_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "xpretrain_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)


def test_state_dict_names_match_reference_layout():
    from types import SimpleNamespace
    from oracle import clipvip_oracle as O
    from xpretrain_b200.modeling import VidCLIP
    add = SimpleNamespace(type="ViP", temporal_size=12, if_use_temporal_embed=1, logit_scale_init_value=4.6, add_cls_num=3)
    from xpretrain_b200.modeling.clip_vip import ClipVipConfig, TowerConfig
    mc = ClipVipConfig(vision=TowerConfig(768, 12, 2, 3072), text=TowerConfig(512, 8, 2, 2048))
    model = VidCLIP(SimpleNamespace(clip_config=mc, clip_weights="", clip_vision_additional_config=add))
    cfg = O.ClipVipCfg(vision=O.TowerCfg(768, 12, 2, 3072), text=O.TowerCfg(512, 8, 2, 2048))
    sd = O.init_state_dict(cfg)      # keyed like the reference CLIPModel.state_dict() (pinned by make_golden.py)
    own = model.clipmodel.state_dict()
    assert set(own) == set(sd)
    for k in sd:
        assert own[k].shape == sd[k].shape and own[k].dtype == sd[k].dtype, k
    assert abs(float(model.clipmodel.logit_scale) - 4.6) < 1e-6
    # weight-decay grouping of the reference (optimization/utils.py:127) keys on these substrings
    names = [n for n, _ in model.named_parameters()]
    assert any(n.endswith("logit_scale") for n in names) and any("pre_layrnorm" in n for n in names)


def test_wgrad_plan_fills_waves():
    from xpretrain_b200.ops import wgrad_plan
    for n_out, n_in in [(3072, 768), (768, 3072), (2304, 768), (768, 768)]:
        bn, s = wgrad_plan(n_out, n_in, 150784)
        tiles = ((n_out + 255) // 256) * ((n_in + 255) // 256) * s        # CTA pairs: 256 x 256 tiles on 74 clusters
        assert tiles / (-(-tiles // 74) * 74) > 0.9


def test_timesformer_module_has_the_reference_state_dict():
    """config #4 drop-in: parameter names/shapes of hd-vila/src/modeling/timesformer.py:421-455 (no kernel is called)."""
    from oracle import timesformer_oracle as TO
    from xpretrain_b200.modeling.timesformer import TimeSformer

    cfg = TO.TimeSformerCfg(depth=2, num_frames=7, H=10, W=16, embed_dim=128, num_heads=2)
    m = TimeSformer(depth=2, num_frames=7, H=10, W=16, embed_dim=128, num_heads=2)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == TO.param_shapes(cfg)
    m.load_state_dict(TO.init_state_dict(cfg, seed=0), strict=True)
    # reference init quirks (timesformer.py:457-464): temporal_fc of every block but the first starts at zero
    m2 = TimeSformer(depth=2, embed_dim=128, num_heads=2)
    assert float(m2.blocks[1].temporal_fc.weight.abs().sum()) == 0.0
    assert float(m2.blocks[0].temporal_fc.weight.abs().sum()) > 0.0
    import pytest
    import torch
    with pytest.raises(Exception):          # no CPU path
        m(torch.zeros(1, 7, 128, 10, 16))


def test_swin3d_module_has_the_reference_state_dict():
    """config #5 drop-in: parameters and buffers of LF-VILA/src/models/video_encoder.py:450-548 (no kernel is called)."""
    import torch
    from oracle import swin3d_oracle as SO
    from xpretrain_b200.modeling.swin3d import SwinTransformer3D

    cfg = SO.Swin3DCfg()
    m = SwinTransformer3D(patch_norm=True, local_window=8)              # the released VideoEncoder config is the default
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == SO.param_shapes(cfg)
    sd = SO.init_state_dict(cfg, seed=0)
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.layers[2].blocks[0].attn.relative_position_index, SO.rel_pos_index(cfg.window_size[2]))
    assert sum(p.numel() for p in m.parameters()) == 89_229_448


def test_init_weights_statistics_follow_the_reference():
    """CLIPPreTrainedModel._init_weights (CLIP_ViP.py:481-522, factor 1, initializer_range 0.02) + the ViP additions
    (`added_cls ~ N(0,1)` :153, `temporal_embedding = 0` :166): sample std of every tensor family within 3 % of the nominal."""
    import torch
    from xpretrain_b200.modeling.clip_vip import CLIPModel, ClipVipConfig
    torch.manual_seed(0)
    cfg = ClipVipConfig()
    m = CLIPModel(cfg)
    def std(t): return float(t.detach().float().std())
    def close(got, want): assert abs(got - want) < 0.03 * want, (got, want)
    ve, te = m.vision_model.embeddings, m.text_model.embeddings
    close(std(te.token_embedding.weight), 0.02); close(std(te.position_embedding.weight), 0.02)
    close(std(ve.patch_embedding.weight), 0.02); close(std(ve.position_embedding.weight), 0.02)
    assert abs(std(ve.class_embedding) - 768 ** -0.5) < 0.15 * 768 ** -0.5      # 768 samples only: wider band
    assert abs(std(ve.added_cls) - 1.0) < 0.1                           # N(0, 1), 3 x 768 samples
    assert float(ve.temporal_embedding.abs().max()) == 0.0
    for tower, tc in ((m.vision_model, cfg.vision), (m.text_model, cfg.text)):
        in_std = tc.hidden_size ** -0.5 * (2 * tc.num_hidden_layers) ** -0.5
        for layer in (tower.encoder.layers[0], tower.encoder.layers[-1]):
            for lin in (layer.self_attn.q_proj, layer.self_attn.k_proj, layer.self_attn.v_proj, layer.mlp.fc2):
                close(std(lin.weight), in_std)
            close(std(layer.self_attn.out_proj.weight), tc.hidden_size ** -0.5)
            close(std(layer.mlp.fc1.weight), (2 * tc.hidden_size) ** -0.5)
            for lin in (layer.self_attn.q_proj, layer.self_attn.out_proj, layer.mlp.fc1, layer.mlp.fc2):
                assert float(lin.bias.abs().max()) == 0.0
            for ln in (layer.layer_norm1, layer.layer_norm2):
                assert float((ln.weight - 1).abs().max()) == 0.0 and float(ln.bias.abs().max()) == 0.0
    close(std(m.visual_projection.weight), 768 ** -0.5); close(std(m.text_projection.weight), 512 ** -0.5)
    assert abs(float(m.logit_scale) - 4.60) < 1e-6


def test_training_restorer_shaped_checkpoint_round_trips(tmp_path):
    """E2E_TrainingRestorer (CLIP-ViP/src/utils/load_save.py:260-327): `restore.pt` = {'global_step', 'model_state_dict',
    'optim_state_dict'} with every fp32 tensor stored as CPU fp16 (`_to_cpu`, :177-192) and turned back into fp32 on load
    (`_to_cuda`, :159-174).  Our module tree and our AdamW must accept that file unchanged: same keys, same optimizer-state
    layout (`step`, `exp_avg`, `exp_avg_sq`; param_groups with lr / betas / eps / weight_decay / correct_bias)."""
    import torch
    from types import SimpleNamespace
    from xpretrain_b200.modeling import VidCLIP
    from xpretrain_b200.modeling.clip_vip import ClipVipConfig, TowerConfig
    from xpretrain_b200.optimization.adamw import AdamW, build_e2e_optimizer_w_lr_mul

    def to_cpu(state):     # restatement of load_save.py:177-192
        if isinstance(state, torch.Tensor):
            ret = state.cpu()
            return ret.half() if "Float" in state.type() else ret
        if isinstance(state, (list, tuple)):
            return type(state)(to_cpu(t) for t in state)
        if isinstance(state, dict):
            return {n: to_cpu(t) for n, t in state.items()}
        return state

    def to_f32(state):     # load_save.py:159-174 without the .cuda()
        if isinstance(state, torch.Tensor):
            return state.float() if "Half" in state.type() else state
        if isinstance(state, (list, tuple)):
            return type(state)(to_f32(t) for t in state)
        if isinstance(state, dict):
            return {n: to_f32(t) for n, t in state.items()}
        return state

    def build(seed):
        torch.manual_seed(seed)
        add = SimpleNamespace(type="ViP", temporal_size=12, if_use_temporal_embed=1, logit_scale_init_value=4.6, add_cls_num=3)
        mc = ClipVipConfig(vision=TowerConfig(768, 12, 1, 3072), text=TowerConfig(512, 8, 1, 2048))
        model = VidCLIP(SimpleNamespace(clip_config=mc, clip_weights="", clip_vision_additional_config=add))
        opt = AdamW(build_e2e_optimizer_w_lr_mul(list(model.named_parameters()), 1e-4, 0.2), lr=1e-4, betas=(0.9, 0.98))
        return model, opt

    model, opt = build(0)
    g = torch.Generator().manual_seed(3)
    for group in opt.param_groups:                       # optimizer state in the reference layout (adamw.py:64-70)
        for p in group["params"]:
            opt.state[p] = {"step": 17, "exp_avg": torch.randn(p.shape, generator=g) * 1e-3,
                            "exp_avg_sq": torch.rand(p.shape, generator=g) * 1e-6}
    ckpt = {"global_step": 17, "model_state_dict": to_cpu(model.state_dict()), "optim_state_dict": to_cpu(opt.state_dict())}
    assert set(ckpt["optim_state_dict"]["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "correct_bias", "params"}
    path = tmp_path / "restore.pt"
    torch.save(ckpt, path)

    model2, opt2 = build(1)                              # different init: everything must come from the file
    loaded = torch.load(path, weights_only=False)
    model2.load_state_dict(to_f32(loaded["model_state_dict"]))
    opt2.load_state_dict(to_f32(loaded["optim_state_dict"]))
    for (n, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        want = a.half().float() if a.is_floating_point() else a
        assert torch.equal(b, want), n                   # fp16 storage is the reference's choice; the round trip adds nothing
        assert b.dtype == a.dtype
    for g1, g2 in zip(opt.param_groups, opt2.param_groups):
        assert {k: v for k, v in g1.items() if k != "params"} == {k: v for k, v in g2.items() if k != "params"}
        for p1, p2 in zip(g1["params"], g2["params"]):
            s1, s2 = opt.state[p1], opt2.state[p2]
            assert s2["step"] == 17 and s2["exp_avg"].dtype == torch.float32
            assert torch.equal(s2["exp_avg"], s1["exp_avg"].half().float())
            assert torch.equal(s2["exp_avg_sq"], s1["exp_avg_sq"].half().float())
    # load_state_dict_with_mismatch (load_save.py:86-115): shape-mismatched and unknown keys are skipped silently
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["clipmodel.vision_model.embeddings.temporal_embedding"] = torch.zeros(1, 8, 768)
    sd["task_head.weight"] = torch.zeros(3)
    own = model2.state_dict()
    toload = {k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}
    missing, unexpected = model2.load_state_dict(toload, strict=False)
    assert unexpected == [] and missing == ["clipmodel.vision_model.embeddings.temporal_embedding"]
