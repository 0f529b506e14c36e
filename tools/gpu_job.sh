#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm" > $O/r02s_t_gemm.log 2>&1; echo "gemm tests rc=$? $(tail -1 $O/r02s_t_gemm.log)"
timeout 120 python tools/gemm_bench.py > $O/r02s_gemm.log 2>&1
echo "rc=$? $(tail -1 $O/r02s_gemm.log)"
