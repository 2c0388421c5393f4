#!/bin/bash
# tools/gpu_band.sh — parity tests (quick set), then the banded workloads
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "${KEXPR:-not config3 and not config2}" 2>&1 | tail -40 > $O/pytest_gpu.txt; T=${PIPESTATUS[0]}
tail -25 $O/pytest_gpu.txt
[ $T -ne 0 ] && exit 1
for w in banded22 cfg5; do
timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_$w.json 2> $O/bench_$w.err; echo "exit $?" >> $O/bench_$w.err
cat $O/bench_$w.json; tail -3 $O/bench_$w.err
done
