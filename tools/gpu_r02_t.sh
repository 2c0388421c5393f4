#!/bin/bash
# tools/gpu_r02_t.sh — ncu --set full of k_rows_bm on R-MAT scale 20 (source-level counters for the instruction breakdown).
O=gpurun_out; mkdir -p $O
CMD="python bench.py --workload rmat20 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --validate-rows 0 --no-prepare"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rows_bm" -s 1 -c 1 -f -o $O/r02t_rmat20_bm $CMD > $O/r02t_ncu.log 2>&1; tail -2 $O/r02t_ncu.log
