"""One fwd+bwd of the Swin-3D step (config #5) inside a cudaProfilerStart/Stop range, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/profile_swin3d.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import swin3d_oracle as SO  # noqa: E402  (synthetic weights / inputs only)
from xpretrain_b200.modeling.swin3d import SwinTransformer3D  # noqa: E402

dev = torch.device("cuda", 0)
cfg = SO.Swin3DCfg()
model = SwinTransformer3D(patch_norm=True, local_window=8, drop_path_rate=0.0)
model.load_state_dict(SO.init_state_dict(cfg, seed=0))
model = model.to(dev).train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
video = SO.synthetic_video(B, 32, 192, 320, cfg, seed=1).to(dev)


def step():
    for p in model.parameters():
        p.grad = None
    out = model(video)[0]
    (out * out).mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
