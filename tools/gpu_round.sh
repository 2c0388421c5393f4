#!/bin/bash
# tools/gpu_round.sh — one gpurun call: microbench, smoke, GPU parity tests, short benches.  Output in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvidia-smi.txt 2>&1
nproc > $O/host.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> $O/host.txt
( cd tools && timeout 120 ./microbench ) > $O/microbench.txt 2>&1
timeout 600 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke exit $?" >> $O/smoke.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -80 > $O/pytest_gpu.txt
timeout 300 python bench.py --workload small --steps 3 --warmup 3 --validate --cpu-seconds 2 > $O/bench_small.json 2> $O/bench_small.err; echo "exit $?" >> $O/bench_small.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --cpu-seconds 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "exit $?" >> $O/bench_cfg2.err
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "exit $?" >> $O/bench_cfg3.err
timeout 600 python bench.py --steps 3 --warmup 3 --mode twophase --no-cpu-baseline > $O/bench_cfg3_twophase.json 2> $O/bench_cfg3_twophase.err
tail -3 $O/smoke.txt; tail -15 $O/pytest_gpu.txt; cat $O/microbench.txt; cat $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
