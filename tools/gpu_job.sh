#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -2 $O/$name.log; }
XP_GEMM_DEBUG=0 TMO=300 run r02h_gemm_bench python tools/gemm_bench.py
XP_GEMM_DEBUG=1 TMO=300 run r02h_gemm_bench_nostore python tools/gemm_bench.py
