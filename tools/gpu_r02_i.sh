#!/bin/bash
# tools/gpu_r02_i.sh — L2 fetch granularity experiment on config 3: bench line + dram / lts sector counters
# of the sort kernel for cudaLimitMaxL2FetchGranularity = default, 32, 64, 128.
O=gpurun_out; mkdir -p $O
python - <<'PY'
import ctypes
rt = ctypes.CDLL("libcudart.so")
v = ctypes.c_size_t()
print("default cudaLimitMaxL2FetchGranularity:", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value)
PY
CMD="python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --validate-rows 0"
for G in default 32 64 128; do
  if [ $G = default ]; then unset BSPGEMM_L2_FETCH; else export BSPGEMM_L2_FETCH=$G; fi
  timeout 300 $CMD > $O/r02i_bench_$G.json 2> $O/r02i_bench_$G.err; echo "bench $G exit $?"
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for G in default 32 64 128; do
  if [ $G = default ]; then unset BSPGEMM_L2_FETCH; else export BSPGEMM_L2_FETCH=$G; fi
  timeout 600 ncu --metrics $M --clock-control none -k regex:k_fused_sort -s 8 -c 1 --csv --log-file $O/r02i_ncu_$G.csv $CMD > /dev/null 2>&1; echo "ncu $G exit $?"
done
python - <<'PY'
import json, csv
for g in ("default", "32", "64", "128"):
    try:
        d = json.loads(open(f"gpurun_out/r02i_bench_{g}.json").read().strip().splitlines()[-1])
        rows = [r for r in csv.reader(open(f"gpurun_out/r02i_ncu_{g}.csv")) if len(r) > 5 and r[0].isdigit()]
        m = {r[-3]: r[-1] for r in rows}
        print(f"L2 fetch {g:8s} kernel {d['roofline']['kernel_ms']:.4f} ms step {d['ms_per_step']:.4f} | " + " ".join(f"{k.split('.')[0].replace('__','_')}={v}" for k, v in m.items()))
    except Exception as e:
        print(g, "FAILED", e)
PY
