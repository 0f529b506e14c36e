#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "$name rc=$?"; tail -2 $O/$name.log | cut -c1-600; }
TMO=900 run r02v_t_all python -m pytest tests -m gpu -q
TMO=300 run r02v_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run r02v_bench python bench.py --steps 8 --warmup 3
TMO=600 run r02v_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches_step_v2.csv python bench.py --steps 1 --warmup 1
