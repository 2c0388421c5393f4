#!/bin/bash
# tools/gpu_r02_d.sh — sort-kernel iteration: parity tests of the ELL / sort paths, then A/B bench lines on the same box
# (floating-point network, integer network, and the library saved as libbspgemm_base.so when present).
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort or ell or config or prepared or fixture or kats or seeded or edge or out_of_range" > $O/r02d_tests.log 2>&1; echo "tests exit $?"; tail -8 $O/r02d_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $B > $O/r02d_flt.json 2> $O/r02d_flt.err; echo "flt exit $?"
BSPGEMM_SORT_INT=1 timeout 300 $B > $O/r02d_int.json 2> $O/r02d_int.err; echo "int exit $?"
if [ -f binary-spgemm_b200/libbspgemm_base.so ]; then BSPGEMM_LIB=$PWD/binary-spgemm_b200/libbspgemm_base.so timeout 300 $B > $O/r02d_base.json 2> $O/r02d_base.err; echo "base exit $?"; fi
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02d_cfg2.json 2> $O/r02d_cfg2.err; echo "cfg2 exit $?"
python - <<'PY'
import json
for w in ("flt", "int", "base", "cfg2"):
    try:
        d = json.loads(open(f"gpurun_out/r02d_{w}.json").read().strip().splitlines()[-1])
        print(w, "ms/step %.4f kernel %.4f frac %.3f | unprepared %.4f | validated %s | %s" % (
            d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["unprepared"]["ms_per_step"], d["validated"]["ok"], d["roofline"]["kernel"]))
    except Exception as e:
        print(w, "FAILED", e)
PY
tail -3 $O/r02d_flt.err
