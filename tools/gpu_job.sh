#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -3 $O/$name.log | cut -c1-400; }
TMO=600 run r02n_t_attn python -m pytest tests/test_gpu_kernels.py -q -k "vip_attention"
TMO=200 run r02n_attn_bench python tools/attn_bench.py
TMO=900 run r02n_t_parity python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py -q
TMO=900 run r02n_bench python bench.py --steps 6 --warmup 3 --no-eager
