#!/bin/bash
# tools/gpu_r02_g.sh — ncu --set full of the CSR-gather kernels on R-MAT scale 20 (k_fused, k_rows_sort, k_rows_window, k_copy_rows).
O=gpurun_out; mkdir -p $O
CMD="python bench.py --workload rmat20 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 0 --no-prepare"
timeout 600 $CMD > $O/r02g_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_fused|k_rows_sort|k_rows_window|k_copy_rows" -s 45 -c 7 -f -o $O/r02g_rmat20 $CMD > $O/r02g_ncu.log 2>&1
tail -3 $O/r02g_ncu.log; head -c 400 $O/r02g_plain.log
