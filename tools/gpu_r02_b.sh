#!/bin/bash
# tools/gpu_r02_b.sh — smoke, the GPU suite, bench lines (cfg3, cfg2) with validation.
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/r02b_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02b_gputests.log 2>&1; echo "gpu tests exit $?"; tail -15 $O/r02b_gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_cfg3.json 2> $O/r02b_bench_cfg3.err; echo "bench cfg3 exit $?"; tail -3 $O/r02b_bench_cfg3.err
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_cfg2.json 2> $O/r02b_bench_cfg2.err; echo "bench cfg2 exit $?"
python - <<'PY'
import json
for w in ("cfg3", "cfg2"):
    try:
        d = json.loads(open(f"gpurun_out/r02b_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "ms/step %.4f kernel %.4f frac %.3f | unprepared %.4f | e2e %.2f ms | validated %s | cold %.1f ms | launches/step %s cached %s" % (
            d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["unprepared"]["ms_per_step"], d["e2e"]["ms_per_step"],
            d["validated"], d["arena"]["cold_first_call_ms"], d["pipeline"]["launches_per_step"], d["pipeline"]["plan_cached"]))
    except Exception as e:
        print(w, "FAILED", e)
PY
