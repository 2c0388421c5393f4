#!/bin/bash
# tools/gpu_r02_nvtx.sh — the NVTX ranges select launches: ncu --nvtx --nvtx-include "bspgemm.fast/" lists only the replayed products' kernels.
O=gpurun_out; mkdir -p $O
timeout 600 ncu --nvtx --nvtx-include "bspgemm.fast/" --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file $O/r02_nvtx_fast.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1; echo "ncu nvtx exit $?"
grep -c "k_fused_sort_async" $O/r02_nvtx_fast.csv; grep -c "k_build_ell\|k_maxlen" $O/r02_nvtx_fast.csv; tail -3 $O/r02_nvtx_fast.csv | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture or kats or prepared" 2>&1 | tail -1
