#!/bin/bash
# tools/gpu_r02_zd.sh — ncu of the config-3 kernel on a 1/64 shard (where does the fixed ~100 us go?)
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused_sort_async" -s 4 -c 1 -f -o $O/r02zd_shard64 python tools/shard_one.py 64 > $O/r02zd_ncu.log 2>&1; tail -3 $O/r02zd_ncu.log
