#!/bin/bash
# tools/gpu_r02_ze.sh — chain: static block ids for the first five iterations.  Parity of the sort / ELL paths, shard sweep, config 3 / 2 bench lines.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort or ell or config2 or config3 or prepared or fixture or kats or seeded or edge or out_of_range or repeated" > $O/r02ze_tests.log 2>&1; echo "tests exit $?"; tail -3 $O/r02ze_tests.log
timeout 600 python tools/shard_sweep.py > $O/r02ze_shard_sweep.txt 2>&1; cat $O/r02ze_shard_sweep.txt
for W in cfg3 cfg2; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/r02ze_$W.json 2> $O/r02ze_$W.err; echo "$W exit $?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ze_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.4f kernel %.4f frac %.3f | unprepared %.4f | validated %s" % (
            d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["unprepared"]["ms_per_step"], d["validated"]["ok"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
