#!/bin/bash
# The GPU-box job of the current iteration (run as `bash tools/gpu_job.sh` under gpurun); logs go to gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -3 $O/$name.log; }
TMO=600 run r02c_t_kernels python -m pytest tests/test_gpu_kernels.py -q -k "vip_attention or layernorm or nce"
for d in 0 1 2 3; do XP_ATTN_BWD_DEBUG=$d TMO=200 run r02c_attn_bench_dbg$d python tools/attn_bench.py; done
TMO=600 run r02c_ncu_attn_bwd ncu --set full --import-source on --clock-control none -k regex:vip_attn_bwd_tc -c 1 -f -o $O/r02_attn_bwd_tc_v3 python tools/attn_bench.py 16
TMO=900 run r02c_t_parity python -m pytest tests/test_gpu_parity.py -q -s
XP_RESIDUAL_BF16=1 TMO=600 run r02c_bench_bf16res python bench.py --steps 5 --warmup 3 --no-eager
TMO=900 run r02c_bench python bench.py --steps 5 --warmup 3 --no-eager
TMO=600 run r02c_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_step_v1.csv python bench.py --steps 1 --warmup 1 --no-eager
