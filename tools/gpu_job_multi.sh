#!/bin/bash
# Multi-GPU job: configs[3] / configs[4] (TimeSformer / Swin-3D encoders) as data-parallel replicas with gradient averaging.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
TMO=400 run r02_bench_timesformer_n$N $TR bench.py --workload timesformer --gpus $N --steps 8 --warmup 3
TMO=600 run r02_bench_swin3d_n$N $TR bench.py --workload swin3d --gpus $N --steps 5 --warmup 3
TMO=400 run r02_bench_n${N}_final $TR bench.py --gpus $N --steps 6 --warmup 3
