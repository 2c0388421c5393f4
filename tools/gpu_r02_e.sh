#!/bin/bash
# tools/gpu_r02_e.sh — ncu --set full of the cfg3 sort kernel: floating-point network and integer network (same binary).
O=gpurun_out; mkdir -p $O
CMD="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 0"
timeout 600 $CMD > $O/r02e_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_sort -s 8 -c 1 -f -o $O/r02e_flt $CMD > $O/r02e_ncu_flt.log 2>&1
tail -2 $O/r02e_ncu_flt.log
export BSPGEMM_SORT_INT=1
timeout 600 $CMD > $O/r02e_plain_int.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_sort -s 8 -c 1 -f -o $O/r02e_int $CMD > $O/r02e_ncu_int.log 2>&1
tail -2 $O/r02e_ncu_int.log
