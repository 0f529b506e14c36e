#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -3 $O/$name.log; }
for d in 4 5 6 7; do XP_ATTN_BWD_DEBUG=$d TMO=200 run r02d_attn_trace_dbg$d python tools/attn_bench.py; done
