#!/bin/bash
# tools/gpu_r02_zc.sh — config-3 kernel time against the shard size on one GPU (the kernel's fixed cost), default and with fewer warps.
O=gpurun_out; mkdir -p $O
timeout 600 python tools/shard_sweep.py > $O/r02zc_shard_sweep.txt 2>&1; cat $O/r02zc_shard_sweep.txt
BSPGEMM_VERBOSE=1 timeout 300 python tools/shard_sweep.py 2>&1 | grep -m2 "k_fused_sort" 
