#!/bin/bash
# tools/gpu_r02_tiny.sh — k_rows_tiny (rows of <= 64 products): parity of the CSR-gather paths (both pipelines), published workload, cfg4.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or kats or seeded or edge or modes or estimate or capacity or rectangular or repeated or iterated or masked or sprand or bin or power_law or bad_arg or staging" > $O/r02tiny_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02tiny_tests.log
timeout 600 python bench.py --workload pub_n5e6_d5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02tiny_pub.json 2> $O/r02tiny_pub.err; echo "pub exit $?"
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02tiny_cfg4.json 2> $O/r02tiny_cfg4.err; echo "cfg4 exit $?"
python - <<'PY'
import json
for w in ("pub", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02tiny_{w}.json").read().strip().splitlines()[-1]); p = d["pipeline"]
        print(w, "ms/step %.3f | sym %.2f main %.2f num %.2f | launches %d | validated %s | frac %.4f" % (d["ms_per_step"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], p["launches_per_step"], d["validated"]["ok"], d["roofline"]["frac"]))
    except Exception as e: print(w, "FAILED", e)
PY
