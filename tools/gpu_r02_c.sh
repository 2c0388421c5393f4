#!/bin/bash
# tools/gpu_r02_c.sh [N] — N-GPU box: the GPU suite (multi-GPU tests included), torchrun bench at N with validation.
N=${1:-2}
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 1500 python -m pytest tests -m gpu -x -q -rs > $O/r02c_gputests_${N}gpu.log 2>&1; echo "gpu tests exit $?"; tail -12 $O/r02c_gputests_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02c_bench_cfg3_${N}gpu.json 2> $O/r02c_bench_cfg3_${N}gpu.err; echo "bench cfg3 N=$N exit $?"; tail -5 $O/r02c_bench_cfg3_${N}gpu.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02c_bench_cfg3_{N}gpu.json").read().strip().splitlines()[-1])
    print("N", N, "ms/step %.4f kernel %.4f | unprepared %.4f | e2e %.2f ms (per rank %.2f) | validated %s" % (
        d["ms_per_step"], d["roofline"]["kernel_ms"], d["unprepared"]["ms_per_step"], d["e2e"]["ms_per_step"], (d.get("e2e_per_rank") or {}).get("ms_per_step", 0), d["validated"]))
except Exception as e:
    print("FAILED", e)
PY
