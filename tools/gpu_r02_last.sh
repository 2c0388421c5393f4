#!/bin/bash
# tools/gpu_r02_last.sh — last check of the round with the NVTX build: full GPU suite, one bench line, DRAM bytes of the dominant kernels of
# config 2 and config 5 (profiles/traffic.json), NVTX-filtered launch list of the main phase.
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02last_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/r02last_gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02last_cfg3.json 2> $O/r02last_cfg3.err; echo "cfg3 exit $?"
for W in cfg2 cfg5; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_fused_sort|k_band" -s 6 -c 1 --csv --log-file $O/r02last_dram_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1; echo "ncu $W exit $?"
done
timeout 600 ncu --nvtx --nvtx-include "bspgemm/fast/" --metrics gpu__time_duration.sum --clock-control none -c 6 --csv --log-file $O/r02last_nvtx_fast.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1; echo "ncu nvtx exit $?"
python - <<'PY'
import json, csv
d = json.loads(open("gpurun_out/r02last_cfg3.json").read().strip().splitlines()[-1])
print("cfg3 ms/step %.4f kernel %.4f frac %.3f validated %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["validated"]["ok"]))
for w in ("cfg2", "cfg5", ):
    rows = list(csv.reader(open(f"gpurun_out/r02last_dram_{w}.csv")))
    h = [r for r in rows if "Metric Name" in r][0]; mi, vi, ki = h.index("Metric Name"), h.index("Metric Value"), h.index("Kernel Name")
    print(w, [(r[ki][:30], r[mi], r[vi]) for r in rows if len(r) > vi and r[0].isdigit()])
rows = list(csv.reader(open("gpurun_out/r02last_nvtx_fast.csv")))
print("nvtx-filtered launches:", [r[4][:40] for r in rows if len(r) > 5 and r[0].isdigit()][:6])
PY
