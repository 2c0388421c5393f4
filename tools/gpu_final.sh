#!/bin/bash
# tools/gpu_final.sh — end-of-round evidence in one gpurun call: gpu_round.sh (smoke, full GPU suite, cfg2/cfg3 benches, ncu launch
# list + one --set full capture of the fused kernel), then the BASELINE configs 4 and 5 and the reference arm.
bash tools/gpu_round.sh ncu || exit 1
O=gpurun_out
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 exit $?"
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref exit $?"
for f in cfg5 cfg4 reference_arm; do python - $O/bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f value %.4g" % (d.get("ms_per_step",0), d.get("value",0)), (d.get("roofline") or {}).get("kernel"), (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
