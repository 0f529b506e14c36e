#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-500; }
TMO=70 run r02x_bench_timesformer python bench.py --workload timesformer --steps 8 --warmup 3
TMO=90 run r02x_bench_swin3d python bench.py --workload swin3d --steps 5 --warmup 3
