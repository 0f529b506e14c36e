"""PyTorch-eager GPU baseline for the bench workload (BASELINE.md §3 "same-box GPU eager baseline").

The reference itself (/root/reference) does not exist on the GPU box, so this times its pinned restatement
(oracle/clipvip_oracle.py: the same torch ops in the same order as CLIP_ViP.py / loss.py) on the B200 under
torch.autocast(bfloat16) — the reference cannot run `.to(bfloat16)`, autocast is its working
bf16 mode.  fwd + InfoNCE + bwd, CUDA-event timed, B = 64 (falls back to 32 / 16 if eager runs out of memory).
This is a measurement tool (it executes oracle/ on purpose); nothing in the product imports it.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clipvip_oracle as O  # noqa: E402


def run(B, steps=5, warmup=2):
    dev = torch.device("cuda", 0)
    cfg = O.ClipVipCfg()
    sd = {k: (v.to(dev).requires_grad_(True) if v.is_floating_point() else v.to(dev))
          for k, v in O.init_state_dict(cfg, seed=0).items()}
    video, ids, mask = O.synthetic_batch(B, 12, 32, cfg, seed=1234)
    video, ids, mask = video.to(dev), ids.to(dev), mask.to(dev)

    def step():
        for v in sd.values():
            if v.is_floating_point():
                v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = O.clip_vip_forward(sd, video, ids, mask, cfg)
            loss = O.nce_learnable_temp_loss(out["vis_features"].float(), out["text_features"].float(), sd["logit_scale"])
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "pytorch-eager (oracle port of the reference) under bf16 autocast", "batch": B, "ms_per_step": round(ms, 2),
            "pairs_per_s": round(B / ms * 1e3, 2), "max_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}


if __name__ == "__main__":
    for B in (64, 32, 16):
        try:
            print(json.dumps(run(B)), flush=True)
            break
        except torch.OutOfMemoryError:
            print(json.dumps({"batch": B, "error": "out of memory in eager"}), flush=True)
            torch.cuda.empty_cache()
