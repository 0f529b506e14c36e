#!/bin/bash
# The GPU-box job of the current iteration (run as `bash tools/gpu_job.sh` under gpurun); logs go to gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -4 $O/$name.log; }
TMO=600 run r02_t_kernels python -m pytest tests/test_gpu_kernels.py -q -s
for d in 0 1 2 3; do XP_ATTN_BWD_DEBUG=$d TMO=200 run r02_attn_bench_dbg$d python tools/attn_bench.py; done
TMO=600 run r02_ncu_attn_bwd ncu --set full --import-source on --clock-control none -k regex:vip_attn_bwd_tc -c 1 -f -o $O/r02_attn_bwd_tc_v2 python tools/attn_bench.py 16
TMO=900 run r02_t_parity python -m pytest tests/test_gpu_parity.py -q -s
XP_RESIDUAL_BF16=1 TMO=900 run r02_t_parity_bf16res python -m pytest tests/test_gpu_parity.py -q -s -k "full_depth or hidden"
TMO=600 run r02_t_boundary python -m pytest tests/test_gpu_boundary.py tests/test_gpu_optim.py tests/test_gpu_metrics.py -q -s
TMO=300 run r02_smoke python -c "import __graft_entry__ as g; g.smoke()"
XP_NO_OVERLAP=1 XP_RESIDUAL_BF16=1 TMO=600 run r02_bench_nooverlap_bf16res python bench.py --steps 5 --warmup 3 --no-eager
XP_RESIDUAL_BF16=1 TMO=600 run r02_bench_bf16res python bench.py --steps 5 --warmup 3 --no-eager
TMO=900 run r02_bench python bench.py --steps 5 --warmup 3
