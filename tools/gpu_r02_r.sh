#!/bin/bash
# tools/gpu_r02_r.sh — rows_bm.cuh with the register-resident path for rows of up to 16384 products: parity, rmat20 / cfg4.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bin or big_rows or power_law or window or wide or round1 or edge or estimate or capacity or staging" > $O/r02v_tests.log 2>&1; echo "tests exit $?"; tail -5 $O/r02v_tests.log
for W in rmat20 cfg4; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02v_$W.json 2> $O/r02v_$W.err; echo "$W exit $?"; tail -2 $O/r02v_$W.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(f.split("/")[-1], "ms/step %.3f | symbolic %.2f main %.2f numeric %.2f | launches %d | validated %s" % (
            d["ms_per_step"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], p["launches_per_step"], (d["validated"] or {}).get("ok")))
    except Exception as e:
        print(f, "FAILED", e)
PY
