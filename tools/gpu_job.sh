#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -3 $O/$name.log; }
TMO=900 run r02g_t_all python -m pytest tests -m gpu -q -x
TMO=300 run r02g_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run r02g_bench python bench.py --steps 6 --warmup 3
