#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -3 $O/$name.log; }
TMO=600 run r02f_t_kernels python -m pytest tests/test_gpu_kernels.py -q -k "vip_attention"
for d in 0 4 7; do XP_ATTN_BWD_DEBUG=$d TMO=200 run r02f_attn_trace_dbg$d python tools/attn_bench.py; done
