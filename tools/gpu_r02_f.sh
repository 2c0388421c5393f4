#!/bin/bash
# tools/gpu_r02_f.sh — full GPU suite (with the full-size config 4 / 5 tests), then config 4: bench line + ncu launch list.
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $O/r02f_gputests.log 2>&1; echo "gpu tests exit $?"; tail -16 $O/r02f_gputests.log
CMD="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 3000"
timeout 1200 $CMD > $O/r02f_bench_cfg4.json 2> $O/r02f_bench_cfg4.err; echo "cfg4 exit $?"; tail -3 $O/r02f_bench_cfg4.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/r02f_cfg4_launches.csv $CMD > $O/r02f_ncu_cfg4.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_cfg4.json").read().strip().splitlines()[-1])
    print("cfg4 ms/step %.2f | pipeline %s | validated %s | unprepared %.2f" % (d["ms_per_step"], d["pipeline"], d["validated"], d["unprepared"]["ms_per_step"]))
except Exception as e:
    print("FAILED", e)
PY
