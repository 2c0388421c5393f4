#!/bin/bash
# tools/gpu_r02_h.sh — skewed-matrix pipeline: parity tests of the CSR-gather paths, rmat20 and config 4 bench lines.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bin or power or window or wide or rows or seeded or kats or edge or rectangular or capacity or estimate or config4 or i64" > $O/r02h_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r02h_tests.log
for W in rmat20 cfg4; do
timeout 1200 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 3000 > $O/r02h_bench_$W.json 2> $O/r02h_bench_$W.err; echo "$W exit $?"; tail -2 $O/r02h_bench_$W.err
done
python - <<'PY'
import json
for w in ("rmat20", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02h_bench_{w}.json").read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(w, "ms/step %.2f | symbolic %.1f main %.1f numeric %.1f est %.2f launches %d | validated %s" % (d["ms_per_step"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], p["ms_estimate"], p["launches_per_step"], d["validated"]["ok"]))
    except Exception as e:
        print(w, "FAILED", e)
PY
