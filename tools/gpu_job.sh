#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -2 $O/$name.log; }
TMO=600 run r02j_t_kernels python -m pytest tests/test_gpu_kernels.py -q -x
TMO=300 run r02j_gemm_bench python tools/gemm_bench.py
TMO=900 run r02j_t_rest python -m pytest tests/test_gpu_parity.py tests/test_gpu_timesformer.py tests/test_gpu_swin3d.py tests/test_gpu_boundary.py tests/test_gpu_optim.py -q -x
TMO=900 run r02j_bench python bench.py --steps 6 --warmup 3 --no-eager
