#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -2 $O/$name.log | cut -c1-400; }
TMO=900 run r02o_t_all python -m pytest tests -m gpu -q
TMO=200 run r02o_attn_bench python tools/attn_bench.py
TMO=900 run r02o_bench python bench.py --steps 8 --warmup 3
