#!/bin/bash
# tools/gpu_r02_a.sh — round-2 starting point: GPU suite, cfg3 bench line, one ncu --set full capture of the shipped hot kernel.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02a_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/r02a_gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_cfg3.json 2> $O/r02a_bench_cfg3.err; echo "bench exit $?"
CMD="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 $CMD > $O/r02a_plain.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_fused_sort -s 3 -c 1 -f -o $O/r02a_async_cfg3 $CMD > $O/r02a_ncu.log 2>&1
tail -3 $O/r02a_ncu.log; head -c 1500 $O/r02a_bench_cfg3.json
