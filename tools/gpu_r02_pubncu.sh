#!/bin/bash
# tools/gpu_r02_pubncu.sh — ncu --set full of k_rows_warp (count pass) on the published sprand workload.
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_rows_warp" -s 6 -c 2 -f -o $O/r02_pub_rows_warp python bench.py --workload pub_n5e6_d5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --validate-rows 0 > $O/r02_pub_ncu.log 2>&1; tail -2 $O/r02_pub_ncu.log
