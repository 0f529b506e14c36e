#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-600; }
TMO=600 run r02l_t_ln python -m pytest tests/test_gpu_kernels.py -q -k "layernorm"
XP_RESIDUAL_DTYPE=fp16 TMO=900 run r02l_t_parity_fp16res python -m pytest tests/test_gpu_parity.py -q -s -k "full_depth or hidden or cfg1 or depth2"
XP_RESIDUAL_DTYPE=fp16 TMO=600 run r02l_bench_fp16res python bench.py --steps 6 --warmup 3 --no-eager
TMO=600 run r02l_bench_timesformer python bench.py --workload timesformer --steps 8 --warmup 3
TMO=900 run r02l_bench_swin3d python bench.py --workload swin3d --steps 5 --warmup 3
TMO=600 run r02l_bench_ref_timesformer python bench.py --impl reference --workload timesformer --steps 3 --warmup 1
TMO=1500 run r02l_memcheck compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_gather or layernorm_with or uint8 or (vip_attention and 2-1-4-20-2) or (vip_attention and 2-2-2-100-3) or nce_loss_and_grads or gemm"
TMO=600 run r02l_ncu_fc1 ncu --set full --clock-control none -k regex:gemm_pair_kernel -s 24 -c 1 -f -o $O/r02_gemm_fc1_tma python tools/gemm_bench.py
TMO=900 run r02l_bench python bench.py --steps 8 --warmup 3
