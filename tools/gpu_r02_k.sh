#!/bin/bash
# tools/gpu_r02_k.sh — new-feature tests (masked product, iterated products, distributed consumer), the published n5e6_d5
# workload, and the M2/L threshold sweep (BSPGEMM_CAP_M2) on R-MAT scale 20 and 22.
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "masked or iterated or distributed or window or power or bin" > $O/r02k_tests.log 2>&1; echo "tests exit $?"; tail -6 $O/r02k_tests.log
timeout 600 python bench.py --workload pub_n5e6_d5 --steps 10 --warmup 3 --cpu-seconds 8 > $O/r02k_bench_pub.json 2> $O/r02k_bench_pub.err; echo "pub exit $?"; tail -2 $O/r02k_bench_pub.err
for C in 16384 8192 4096; do
  BSPGEMM_CAP_M2=$C timeout 600 python bench.py --workload rmat20 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 3000 > $O/r02k_rmat20_$C.json 2> $O/r02k_rmat20_$C.err; echo "rmat20 cap_m2=$C exit $?"
done
for C in 16384 4096; do
  BSPGEMM_CAP_M2=$C timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 3000 > $O/r02k_cfg4_$C.json 2> $O/r02k_cfg4_$C.err; echo "cfg4 cap_m2=$C exit $?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02k_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d["pipeline"]
        print(f.split("/")[-1], "ms/step %.3f | symbolic %.2f main %.2f numeric %.2f | rows s/m/l %d/%d/%d | validated %s | e2e %.1f | cpu %s" % (
            d["ms_per_step"], p["ms_symbolic"], p["ms_main"], p["ms_numeric"], p["rows_s"], p["rows_m"], p["rows_l"], d["validated"]["ok"], d["e2e"]["ms_per_step"],
            (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "FAILED", e)
PY
