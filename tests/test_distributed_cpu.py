"""CPU, world_size 2 over gloo: the N > 1 host logic (rank-major differentiable all-gather with a collective-free
backward, gradient averaging) against the in-repo pinned semantics of the reference
(LF-VILA/src/utils/dist.py:21-41 SyncFunction: all_gather forward, all_reduce(SUM)+slice backward)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sync_function_reference(t):
    """The reference's SyncFunction, restated: gather forward; SUM-all-reduce then local slice backward."""
    class Ref(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.b = x.shape[0]
            out = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
            dist.all_gather(out, x)
            return torch.cat(out, 0)

        @staticmethod
        def backward(ctx, g):
            g = g.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            r = dist.get_rank()
            return g[r * ctx.b:(r + 1) * ctx.b]
    return Ref.apply(t)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from oracle import clipvip_oracle as O
    from xpretrain_b200.utils import distributed as xd
    r, _, w = xd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)
    b, d = 4, 16
    vis = torch.nn.functional.normalize(torch.randn(b, d), dim=-1)
    txt = torch.nn.functional.normalize(torch.randn(b, d), dim=-1)
    temp = torch.tensor(2.0)
    res = {}
    for name, gather in (("ours", xd.allgather), ("ref", _sync_function_reference)):
        v, t, p = vis.clone().requires_grad_(True), txt.clone().requires_grad_(True), temp.clone().requires_grad_(True)
        V, T = gather(v), gather(t)
        loss = O.nce_learnable_temp_loss(V, T, p)
        loss.backward()
        res[name] = (V.detach(), loss.detach(), v.grad, t.grad, p.grad)
    ok = True
    for a, c in zip(res["ours"], res["ref"]):
        ok = ok and torch.allclose(a, c, atol=1e-6)
    # rank-major order: rows [r*b, (r+1)*b) of the gathered matrix are rank r's
    ok = ok and torch.equal(res["ours"][0][rank * b:(rank + 1) * b], vis)
    # gradient averaging
    lin = torch.nn.Linear(3, 2)
    for p_ in lin.parameters():
        p_.grad = torch.full_like(p_, float(rank + 1))
    xd.average_gradients(lin.parameters(), bucket_bytes=8)
    ok = ok and all(torch.allclose(p_.grad, torch.full_like(p_, 1.5)) for p_ in lin.parameters())
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allgather_and_grad_average_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, True), (1, True)]
