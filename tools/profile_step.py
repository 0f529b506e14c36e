"""One training step of the bench workload inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_step.py
(launch list: compare kernel SHARES of the step, the absolute times are cold-cache and serialised), or for
    ncu --profile-from-start off --set full -k regex:gemm_kernel -c 4 -o ... python tools/profile_step.py --fwd-only
"""
import argparse
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from xpretrain_b200.modeling import VidCLIP  # noqa: E402
from xpretrain_b200.optimization.loss import gather_nce_loss  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--fwd-only", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
add = SimpleNamespace(type="ViP", temporal_size=12, if_use_temporal_embed=1, logit_scale_init_value=4.6, add_cls_num=3)
torch.manual_seed(0)
model = VidCLIP(SimpleNamespace(clip_config="openai/clip-vit-base-patch16", clip_weights="",
                                clip_vision_additional_config=add)).to(dev)
B = args.batch
video = torch.randn(B, 12, 3, 224, 224, device=dev)
ids = torch.randint(1, 49406, (B, 32), device=dev)
ids[:, -1] = 49407
mask = torch.ones(B, 32, dtype=torch.long, device=dev)


def step():
    for p in model.parameters():
        p.grad = None
    out = model(video=video, text_input_ids=ids, text_input_mask=mask)
    loss = gather_nce_loss(out["vis_features"], out["text_features"], model.clipmodel.logit_scale)
    if not args.fwd_only:
        loss.backward()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
