#!/bin/bash
# tools/gpu_r02_multi.sh — N-GPU runs (gpurun --gpus N -- 'N=<N> bash tools/gpu_r02_multi.sh'): the >= 2-GPU tests, then the
# bench lines of config 3, config 5 and config 4 (equal rows and equal intermediate products per rank), validated per rank.
O=gpurun_out; mkdir -p $O
N=${N:-2}
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or over_nccl or sharded or consumer or validity_driver or reference_main" > $O/r02f${N}_gputests.log 2>&1; echo "gpu tests exit $?"; tail -4 $O/r02f${N}_gputests.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r02f${N}_cfg3.json 2> $O/r02f${N}_cfg3.err; echo "cfg3 exit $?"
timeout 900 $TR bench.py --gpus $N --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02f${N}_cfg5.json 2> $O/r02f${N}_cfg5.err; echo "cfg5 exit $?"
[ -n "$ROWS" ] && { timeout 900 $TR bench.py --gpus $N --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02f${N}_cfg4_rows.json 2> $O/r02f${N}_cfg4_rows.err; echo "cfg4 rows exit $?"; }
timeout 900 $TR bench.py --gpus $N --workload cfg4 --split ip --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02f${N}_cfg4_ip.json 2> $O/r02f${N}_cfg4_ip.err; echo "cfg4 ip exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f${N}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f.split("/")[-1], "gpus %d ms/step %.3f kernel %.3f frac %.3f | validated %s | e2e %s" % (
            d["n_gpus"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], (d["validated"] or {}).get("ok"), ("%.1f ms" % e["ms_per_step"]) if e else None))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
