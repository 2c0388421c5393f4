#!/bin/bash
# tools/gpu_prof_rmat.sh — ncu of the windowed-bitmap kernels on the R-MAT scale-20 workload (after a plain run exits 0)
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --workload rmat20 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$CMD > $O/plain_rmat20.log 2>&1 || { tail -5 $O/plain_rmat20.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_rmat20.csv $CMD > $O/ncu_launches_rmat20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rows_window -s 18 -c 6 -o $O/prof_window $CMD > $O/ncu_window.log 2>&1
tail -3 $O/ncu_window.log
