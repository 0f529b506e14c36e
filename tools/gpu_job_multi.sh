#!/bin/bash
# Multi-GPU job: 2-rank NCCL tests (fused exchange + InfoNCE, overlapped gradient averaging) and the N-rank bench line.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TMO=300 run r02w_t_multirank python -m pytest tests/test_gpu_multirank.py -m gpu -q
TMO=400 run r02_bench_n${N}_v8 $TR bench.py --gpus $N --steps 6 --warmup 3
