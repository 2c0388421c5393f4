#!/bin/bash
# tools/gpu_r02_cap.sh — cheap-row matrices: warp-bin capacity from the measured maximum.  Parity of the CSR-gather paths, published workload, config 3 unchanged.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or kats or seeded or edge or modes or estimate or capacity or rectangular or repeated or iterated or masked" > $O/r02cap_tests.log 2>&1; echo "tests exit $?"; tail -2 $O/r02cap_tests.log
timeout 600 python bench.py --workload pub_n5e6_d5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02cap_pub.json 2> $O/r02cap_pub.err; echo "pub exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02cap_pub.json").read().strip().splitlines()[-1]); p = d["pipeline"]
print("pub ms/step %.3f | cap_s %d | validated %s | e2e %.1f ms | frac %.4f" % (d["ms_per_step"], p["cap_s"], d["validated"]["ok"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
PY
