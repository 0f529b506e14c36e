"""CPU: host-side logic of the widened rows — schedules / parameter grouping, attention sequence descriptors, Swin-3D index
tables — checked against the oracles (no kernel is launched)."""
import os

import pytest
import torch
import torch.nn.functional as F


def test_lr_schedule_and_grouping_match_oracle_and_reference_golden(golden_dir):
    from oracle import adamw_oracle as AO
    from xpretrain_b200.optimization.adamw import build_e2e_optimizer_w_lr_mul, get_lr_sched

    gold = torch.load(os.path.join(golden_dir, "adamw_8steps.pt"), weights_only=False)
    cfg = gold["cfg"]
    for step in range(1, cfg["steps"] + 1):          # the lrs in the golden came from the reference's get_lr_sched
        assert get_lr_sched(step, cfg["decay"], cfg["learning_rate"], cfg["num_train_steps"], warmup_ratio=cfg["warmup_ratio"]) \
            == gold["lrs"][step - 1]
    for decay in ("linear", "cosine", "invsqrt", "constant"):
        for step in (0, 1, 3, 4, 5, 17, 20, 25):
            assert get_lr_sched(step, decay, 2e-4, 20, warmup_ratio=0.2) == AO.lr_schedule(step, decay, 2e-4, 20, 0.2)
    named = [(n, torch.nn.Parameter(torch.zeros(s))) for n, s in gold["shapes"].items()]
    name_of = {id(p): n for n, p in named}
    groups = build_e2e_optimizer_w_lr_mul(named, cfg["learning_rate"], cfg["weight_decay"], lr_mul=cfg["lr_mul"],
                                          lr_mul_prefix=cfg["lr_mul_prefix"])
    assert [[name_of[id(p)] for p in g["params"]] for g in groups] == gold["group_names"]
    assert [g["weight_decay"] for g in groups] == [cfg["weight_decay"], 0.0, cfg["weight_decay"], 0.0]
    assert groups[0]["lr"] == cfg["lr_mul"] * cfg["learning_rate"] and "lr" not in groups[2]


def _desc_rows(d):
    """Expand an XpSegAttn stride descriptor into (sequence, position) -> row and the segment id of every position."""
    seqs = []
    for s in range(d.n_seq):
        base = (s // d.inner) * d.outer_stride + (s % d.inner) * d.inner_stride
        if base >= d.n_rows:
            continue
        fit = (d.n_rows - base + d.tok_stride - 1) // d.tok_stride
        n = min(d.seq_len, fit)
        seqs.append([(base + i * d.tok_stride, i // d.seg_len) for i in range(n)])
    return seqs


@pytest.mark.parametrize("B,T,H,W", [(2, 7, 2, 5), (1, 8, 3, 3), (3, 3, 1, 5)])
def test_timesformer_descriptors_cover_the_einops_groups(B, T, H, W):
    """temporal_desc / spatial_desc (stride patterns over '(h w t)'-ordered rows) == the reference's rearranges
    'b (h w t) m -> (b h w) t m' and '-> (b t) (h w) m' (timesformer.py:210,217)."""
    from einops import rearrange
    from xpretrain_b200 import ops

    HW = H * W
    rows = torch.arange(B * HW * T).view(B, HW * T, 1)
    want_t = {tuple(g.tolist()) for g in rearrange(rows, 'b (h w t) m -> (b h w) (t m)', h=H, w=W, t=T)}
    want_s = {tuple(g.tolist()) for g in rearrange(rows, 'b (h w t) m -> (b t) (h w m)', h=H, w=W, t=T)}
    got_t = set()
    for seq in _desc_rows(ops.temporal_desc(B * HW * T, T, 2, 384, 128)):
        by_seg = {}
        for row, seg in seq:
            by_seg.setdefault(seg, []).append(row)
        got_t |= {tuple(v) for v in by_seg.values()}
    assert got_t == want_t
    got_s = {tuple(r for r, _ in seq) for seq in _desc_rows(ops.spatial_desc(B, T, HW, 2, 384, 128))}
    assert got_s == want_s
    assert all(seg == 0 for seq in _desc_rows(ops.spatial_desc(B, T, HW, 2, 384, 128)) for _, seg in seq)   # dense


@pytest.mark.parametrize("B,D,H,W,layer", [(1, 4, 7, 7, 0), (2, 8, 6, 10, 0), (1, 8, 3, 5, 1), (2, 8, 2, 3, 2)])
def test_swin3d_index_tables_equal_the_reference_pad_roll_partition(B, D, H, W, layer):
    """The window index tables (plain and shifted) and the shift mask of modeling/swin3d.py vs the oracle's
    F.pad / torch.roll / window_partition applied to a tagged tensor (video_encoder.py:214-230, 309-322)."""
    from oracle import swin3d_oracle as SO
    from xpretrain_b200.modeling import swin3d as S

    m = S.SwinTransformer3D(embed_dim=64, depths=[2, 2, 2], num_heads=[2, 4, 8], stages=[0, 1, 2], downsample_stages=[0, 1],
                            window_size=[[2, 3, 5], [4, 3, 5], [8, 3, 5]], patch_norm=True, local_window=8)
    geo = S._layer_geometry(m, layer, B, D, H, W, torch.device("cpu"))
    ws, ss = geo["ws"], geo["ss"]
    x = torch.arange(B * D * H * W, dtype=torch.float32).view(B, D, H, W, 1) + 1        # row + 1; padding will be 0
    xp = F.pad(x, (0, 0, 0, (ws[2] - W % ws[2]) % ws[2], 0, (ws[1] - H % ws[1]) % ws[1], 0, (ws[0] - D % ws[0]) % ws[0]))
    assert geo["n_pad"] == int((xp == 0).sum()) and geo["L"] == ws[0] * ws[1] * ws[2]
    for which, t in ((0, xp), (1, torch.roll(xp, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3)))):
        want = SO.window_partition(t, ws).squeeze(-1).long() - 1
        got = geo["idx"][which].long()
        real = want >= 0
        assert torch.equal(got[real], want[real])
        pads = got[~real]
        assert bool((pads >= geo["n_real"]).all()) and pads.unique().numel() == pads.numel()   # each pad slot has its own row
    if any(s > 0 for s in ss):
        assert torch.equal(geo["mask"], SO.shift_mask(xp.shape[1], xp.shape[2], xp.shape[3], ws, ss))
    else:
        assert geo["mask"] is None
    # patch merging: the 2x2 gather table against the oracle's strided-slice concat on the tagged tensor
    midx, H2, W2 = S._merge_index(m, B, D, H, W, torch.device("cpu"))
    xm = F.pad(x, (0, 0, 0, W % 2, 0, H % 2)) if (H % 2 or W % 2) else x
    want = torch.cat([xm[:, :, 0::2, 0::2], xm[:, :, 1::2, 0::2], xm[:, :, 0::2, 1::2], xm[:, :, 1::2, 1::2]], -1).reshape(-1).long() - 1
    assert (H2, W2) == ((H + 1) // 2, (W + 1) // 2) and torch.equal(midx.long(), want)


def test_retrieval_metrics_host_part_matches_reference_semantics_with_ties():
    """utils.metrics.metrics_from_counts (pure numpy) vs the oracle's compute_metrics on matrices with many exact ties."""
    import numpy as np
    from oracle import metrics_oracle as MO
    from xpretrain_b200.utils.metrics import metrics_from_counts

    rng = np.random.RandomState(0)
    for n in (1, 2, 7, 40):
        for levels in (3, 1000):                                 # few distinct values -> lots of ties with the diagonal
            x = rng.randint(0, levels, size=(n, n)).astype(np.float32)
            g, e = MO.rank_counts(x)
            assert tuple(float(v) for v in metrics_from_counts(g, e)) == tuple(float(v) for v in MO.compute_metrics(x))
            # and the oracle's count-based form equals the reference's sort-based definition
            sx = np.sort(-x, axis=1)
            ind = np.where(sx - np.diag(-x)[:, None] == 0)[1]
            assert np.array_equal(np.sort(ind), np.sort(MO.ranks_from_counts(g, e)))
