#!/bin/bash
# tools/gpu_prof.sh [workload] — plain bench run, then (if it exited 0) one ncu --set full capture of the fused kernel.
W=${1:-cfg3}
O=gpurun_out; mkdir -p $O
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 $CMD > $O/plain_$W.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o $O/prof_$W $CMD > $O/ncu_$W.log 2>&1
tail -3 $O/ncu_$W.log; head -c 600 $O/plain_$W.log
