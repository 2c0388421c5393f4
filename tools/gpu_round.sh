#!/bin/bash
# tools/gpu_round.sh — one gpurun call: smoke, GPU parity tests, short benches, then the ncu passes.  Output in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/*.json $O/*.err $O/*.txt
nproc > $O/host.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> $O/host.txt
timeout 600 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; echo "smoke exit $?" >> $O/smoke.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short 2>&1 | tail -80 > $O/pytest_gpu.txt
timeout 300 python bench.py --workload small --steps 3 --warmup 3 --validate --cpu-seconds 2 > $O/bench_small.json 2> $O/bench_small.err; echo "exit $?" >> $O/bench_small.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --cpu-seconds 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "exit $?" >> $O/bench_cfg2.err
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "exit $?" >> $O/bench_cfg3.err
tail -3 $O/smoke.txt; tail -15 $O/pytest_gpu.txt; cat $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
if [ "$1" == "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
  $CMD > $O/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
  $CMD > $O/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o $O/prof_fused $CMD > $O/ncu_full.log 2>&1
  tail -5 $O/ncu_full.log
fi
