#!/bin/bash
# tools/gpu_r02_y.sh — polled completion wait + bound call: config 3 / config 2 bench lines (step minus kernel = the per-call fixed cost), quick parity.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prepared or fixture or kats or repeated or config2" > $O/r02y_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02y_tests.log
for W in cfg3 cfg2; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/r02y_$W.json 2> $O/r02y_$W.err; echo "$W exit $?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02y_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.4f kernel %.4f gap %.1f us frac %.3f | unprepared %.4f | validated %s" % (
            d["ms_per_step"], d["roofline"]["kernel_ms"], 1e3 * (d["ms_per_step"] - d["roofline"]["kernel_ms"]), d["roofline"]["frac"], d["unprepared"]["ms_per_step"], d["validated"]["ok"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
