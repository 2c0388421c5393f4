#!/bin/bash
# tools/gpu_round.sh [ncu] [quick] — one gpurun call: smoke, GPU parity tests, short benches, then (only if everything
# before exited 0) the ncu passes.  Output in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/*.json $O/*.err $O/*.txt
nproc > $O/host.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> $O/host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv >> $O/host.txt 2>&1
timeout 600 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; S=$?; echo "smoke exit $S" >> $O/smoke.txt
tail -3 $O/smoke.txt
[ $S -ne 0 ] && { tail -30 $O/smoke.txt; exit 1; }
if [[ " $* " == *" quick "* ]]; then
  timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x -k "not config3 and not config2" 2>&1 | tail -40 > $O/pytest_gpu.txt; T=${PIPESTATUS[0]}
else
  timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -60 > $O/pytest_gpu.txt; T=${PIPESTATUS[0]}
fi
tail -25 $O/pytest_gpu.txt
[ $T -ne 0 ] && exit 1
timeout 300 python bench.py --workload small --steps 3 --warmup 3 --validate --cpu-seconds 2 > $O/bench_small.json 2> $O/bench_small.err; echo "exit $?" >> $O/bench_small.err
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --cpu-seconds 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "exit $?" >> $O/bench_cfg2.err
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; B=$?; echo "exit $B" >> $O/bench_cfg3.err
cat $O/bench_cfg2.json; cat $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
[ $B -ne 0 ] && exit 1
if [[ " $* " == *" ncu "* ]]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
  $CMD > $O/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
  $CMD > $O/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o $O/prof_fused $CMD > $O/ncu_full.log 2>&1
  tail -5 $O/ncu_full.log
fi
