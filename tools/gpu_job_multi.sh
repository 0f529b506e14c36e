#!/bin/bash
# Multi-GPU job (run as `bash tools/gpu_job_multi.sh N` under `gpurun --gpus N`): the 2-rank parity test of the fused
# gather + NCE kernel and the overlapped gradient averaging, then the N-rank bench with and without the SM reservation.
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -2 $O/$name.log | cut -c1-400; }
nvidia-smi topo -m > $O/r02_topo_n$N.log 2>&1
TMO=600 run r02_t_multirank_n$N python -m pytest tests/test_gpu_multirank.py -q -s
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
XP_SM_RESERVE=4 TMO=600 run r02_bench_n${N}_reserve4 $TR bench.py --gpus $N --steps 6 --warmup 3
XP_SM_RESERVE=0 TMO=600 run r02_bench_n${N}_reserve0 $TR bench.py --gpus $N --steps 6 --warmup 3
if [ "$N" = "8" ]; then
  TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
  TMO=600 run r02_bench_n4_reserve4 $TR4 bench.py --gpus 4 --steps 6 --warmup 3
fi
