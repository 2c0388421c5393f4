#!/bin/bash
# tools/gpu_rmat.sh — one gpurun call for the CSR-gather (power-law) path: parity tests of the bins, then R-MAT benches.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "${KEXPR:-not config3 and not config2}" 2>&1 | tail -40 > $O/pytest_gpu.txt; T=${PIPESTATUS[0]}
tail -25 $O/pytest_gpu.txt
[ $T -ne 0 ] && exit 1
timeout 300 python bench.py --workload rmat20 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_rmat20.json 2> $O/bench_rmat20.err; echo "exit $?" >> $O/bench_rmat20.err
cat $O/bench_rmat20.json; tail -3 $O/bench_rmat20.err
timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "exit $?" >> $O/bench_cfg4.err
cat $O/bench_cfg4.json; tail -3 $O/bench_cfg4.err
