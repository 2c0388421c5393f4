#!/bin/bash
# tools/gpu_r02_z.sh — rows_bm.cuh compressed single pass for rows of up to 16384 products: parity (incl. the fallback), cfg4 with and without it, rmat20.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bin or big_rows or power_law or window or wide or round1 or edge or estimate or capacity or staging or bad_arg" > $O/r02za_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02za_tests.log
for W in cfg4 rmat20; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02za_$W.json 2> $O/r02za_$W.err; echo "$W exit $?"; tail -2 $O/r02za_$W.err
done
BSPGEMM_BM_NO_COMP=1 timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 0 > $O/r02za_cfg4_nocomp.json 2> /dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02za_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(f.split("/")[-1], "ms/step %.3f | est %.2f symbolic %.2f main %.2f numeric %.2f | validated %s | frac %.4f" % (
            d["ms_per_step"], p["ms_estimate"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], (d["validated"] or {}).get("ok"), d["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
