#!/bin/bash
# tools/gpu_r02_l.sh — L2-bitmap kernel for the big rows + unordered passes for cheap rows: parity, then rmat20 / cfg4 / published workload.
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not config3 and not config2 and not config5 and not coo2csc" > $O/r02l_tests.log 2>&1; echo "tests exit $?"; tail -6 $O/r02l_tests.log
for W in rmat20 cfg4 pub_n5e6_d5; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 6000 > $O/r02l_$W.json 2> $O/r02l_$W.err; echo "$W exit $?"; tail -2 $O/r02l_$W.err
done
BSPGEMM_NO_L2BM=1 timeout 600 python bench.py --workload rmat20 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 0 > $O/r02l_rmat20_nol2bm.json 2> /dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02l_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(f.split("/")[-1], "ms/step %.3f | symbolic %.2f main %.2f numeric %.2f | launches %d | validated %s | e2e %.1f" % (
            d["ms_per_step"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], p["launches_per_step"], (d["validated"] or {}).get("ok"), d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
