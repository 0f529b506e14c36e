#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-700; }
XP_GEMM_NO_TMA_AUX=1 TMO=300 run r02k_gemm_bench_aux_direct python tools/gemm_bench.py
