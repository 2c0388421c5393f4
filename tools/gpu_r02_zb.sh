#!/bin/bash
# tools/gpu_r02_zb.sh — k_rows_sort with the next row prefetched: parity + cfg4 / rmat20; ncu of the long-row kernel k_rows_bm<STAGE,false> (rmat20).
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bin or big_rows or power_law or window or wide or round1 or edge or estimate or capacity or staging" > $O/r02zb_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02zb_tests.log
for W in cfg4 rmat20; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02zb_$W.json 2> $O/r02zb_$W.err; echo "$W exit $?"; tail -2 $O/r02zb_$W.err
done
CMD="python bench.py --workload rmat20 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --validate-rows 0 --no-prepare"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rows_bm" -s 2 -c 2 -f -o $O/r02zb_rmat20_bm $CMD > $O/r02zb_ncu.log 2>&1; tail -2 $O/r02zb_ncu.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02zb_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(f.split("/")[-1], "ms/step %.3f | est %.2f symbolic %.2f main %.2f numeric %.2f | validated %s | frac %.4f" % (
            d["ms_per_step"], p["ms_estimate"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], (d["validated"] or {}).get("ok"), d["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
