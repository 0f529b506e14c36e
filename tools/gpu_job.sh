#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -1 $O/$name.log | cut -c1-500; }
TMO=900 run r02m_bench_swin3d python bench.py --workload swin3d --steps 5 --warmup 3
TMO=900 run r02m_t_all python -m pytest tests -m gpu -q
TMO=300 run r02m_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run r02m_bench python bench.py --steps 8 --warmup 3
TMO=600 run r02m_bench_ref python bench.py --impl reference --steps 8 --warmup 3
