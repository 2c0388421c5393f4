#!/bin/bash
# tools/gpu_r02_w.sh — rows_bm.cuh micro-optimisations (branch-free emit loop, uniform slot count): parity, rmat20 / cfg4, cfg4 with the S bin up to 1024 products.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bin or big_rows or power_law or window or wide or round1 or edge or estimate or capacity or staging" > $O/r02w_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02w_tests.log
for W in rmat20 cfg4; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02w_$W.json 2> $O/r02w_$W.err; echo "$W exit $?"; tail -2 $O/r02w_$W.err
done
BSPGEMM_CAP_S=1024 timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02w_cfg4_caps1024.json 2> /dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02w_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(f.split("/")[-1], "ms/step %.3f | est %.2f symbolic %.2f main %.2f numeric %.2f | launches %d | validated %s" % (
            d["ms_per_step"], p["ms_estimate"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], p["launches_per_step"], (d["validated"] or {}).get("ok")))
    except Exception as e:
        print(f, "FAILED", e)
PY
