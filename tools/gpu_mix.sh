#!/bin/bash
# tools/gpu_mix.sh — quick parity set, then banded + R-MAT benches, then ncu of k_band and of the big-row kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "${KEXPR:-not config3 and not config2}" 2>&1 | tail -40 > $O/pytest_gpu.txt; T=${PIPESTATUS[0]}
tail -25 $O/pytest_gpu.txt
[ $T -ne 0 ] && exit 1
for w in cfg5 rmat20 cfg4; do
timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_$w.json 2> $O/bench_$w.err; echo "exit $?" >> $O/bench_$w.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_$w.json")); print("$w", round(d["ms_per_step"],3), "ms", d["pipeline"])
except Exception as e: print("$w failed", e)
PY
tail -2 $O/bench_$w.err
done
if [[ " $* " == *" ncu "* ]]; then
  CMD="python bench.py --workload banded22 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
  ncu --set full --clock-control none --import-source on -k regex:k_band -s 3 -c 1 -o $O/prof_band $CMD > $O/ncu_band.log 2>&1
  CMD="python bench.py --workload rmat20 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_rmat20.csv $CMD > $O/ncu_launches_rmat20.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"k_rows_sort|k_rows_window|k_fused|k_copy_rows" -s 21 -c 7 -o $O/prof_bigrows $CMD > $O/ncu_bigrows.log 2>&1
  tail -2 $O/ncu_bigrows.log
fi
