#!/bin/bash
# Multi-GPU tuning job: N-rank bench with different SM reservations for NCCL and gradient all-reduce dtypes.
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-300; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
XP_SM_RESERVE=8 TMO=400 run r02_bench_n${N}_reserve8 $TR bench.py --gpus $N --steps 6 --warmup 3
XP_SM_RESERVE=0 XP_GRAD_COMM=bf16 TMO=400 run r02_bench_n${N}_reserve0_bf16comm $TR bench.py --gpus $N --steps 6 --warmup 3
XP_SM_RESERVE=0 TMO=400 run r02_bench_n${N}_reserve0_b $TR bench.py --gpus $N --steps 6 --warmup 3
