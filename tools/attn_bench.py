"""CUDA-event timing of the ViP attention kernels at the bench shape (T=12, H=12, L=196, M=4)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xpretrain_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, T, L, M = 12, 12, 196, 4
C, S = 64 * H, M + T * L
dev = torch.device("cuda", 0)
qkv = (torch.randn(B * S, 3 * C, device=dev) * 0.5).to(torch.bfloat16)
out = torch.empty(B * S, C, dtype=torch.bfloat16, device=dev)
dout = torch.randn(B * S, C, device=dev).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, S, device=dev)
delta = torch.empty(B, H, S, device=dev)
ws = ops.vip_attention_workspace(B, H, T, M, dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fl_f = B * H * (2 * 2 * T * L * (M + L) * 64 + 2 * 2 * M * S * 64)
cases = [("fwd mma.sync", lambda: ops.vip_attention_fwd(qkv, out, lse, ws, B, H, T, L, M, C), fl_f),
         ("bwd mma.sync", lambda: ops.vip_attention_bwd(qkv, out, dout, lse, dqkv, ws, B, H, T, L, M, C, 0.125), 2.5 * fl_f)]
if hasattr(ops, "vip_attention_fwd_tc"):
    cases.insert(1, ("fwd tcgen05", lambda: ops.vip_attention_fwd_tc(qkv, out, lse, ws, B, H, T, L, M, C), fl_f))
if hasattr(ops, "vip_attention_bwd_tc"):
    cases.append(("bwd tcgen05", lambda: ops.vip_attention_bwd_tc(qkv, out, dout, lse, dqkv, ws, delta, B, H, T, L, M, C, 0.125), 2.5 * fl_f))
for name, fn, fl in cases:
    ms = timeit(fn)
    print(f"{name:14s} B={B}: {ms:8.3f} ms   {fl / ms / 1e9:8.1f} TFLOP/s (algorithmic)   -> {ms * 64 / B:7.3f} ms per layer at B=64")
