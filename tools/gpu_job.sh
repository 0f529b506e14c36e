#!/bin/bash
# The GPU-box job of the current iteration (run as `bash tools/gpu_job.sh` under gpurun); logs go to gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -4 $O/$name.log; }
TMO=600 run r02_t_kernels python -m pytest tests/test_gpu_kernels.py -q -x -s
TMO=300 run r02_attn_bench python tools/attn_bench.py
TMO=900 run r02_t_parity python -m pytest tests/test_gpu_parity.py -q -s
TMO=600 run r02_t_boundary python -m pytest tests/test_gpu_boundary.py tests/test_gpu_optim.py tests/test_gpu_metrics.py -q -s
TMO=900 run r02_t_enc python -m pytest tests/test_gpu_timesformer.py tests/test_gpu_swin3d.py -q
TMO=300 run r02_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run r02_bench python bench.py --steps 5 --warmup 3
