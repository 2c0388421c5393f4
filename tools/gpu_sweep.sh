#!/bin/bash
# tools/gpu_sweep.sh "<ENV=val ...>" ... — one short cfg3/cfg2 bench per environment setting (tuning experiments)
O=gpurun_out; mkdir -p $O; : > $O/sweep.txt
for envs in "$@"; do
  for w in cfg2 cfg3; do
    env $envs timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 > $O/sweep_tmp.json
    python - "$envs" $w <<'PY' >> $O/sweep.txt
import json,sys
try:
    d=json.loads(open("gpurun_out/sweep_tmp.json").read())
    print("%-40s %s ms/step %.3f kernel %.3f frac %.3f %s" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["kernel"]))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
done
cat $O/sweep.txt
