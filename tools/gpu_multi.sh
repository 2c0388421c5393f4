#!/bin/bash
# tools/gpu_multi.sh N — multi-GPU checks on one box: validity driver over N GPUs (NCCL broadcast of B inside libbspgemm),
# torchrun bench at N ranks (small workload validated against the oracle, then config 3), reference arm.
N=${1:-2}
O=gpurun_out; mkdir -p $O
python __graft_entry__.py > $O/build.txt 2>&1
( cd binary-spgemm_b200 && BSPGEMM_GPUS=$N make test ) > $O/make_test_$N.txt 2>&1; echo "make test exit $?" >> $O/make_test_$N.txt; tail -3 $O/make_test_$N.txt
( cd binary-spgemm_b200 && BSPGEMM_GPUS=$N host/SpGEMM_gpu ../oracle/_ref/validity_test.mtx 6250 2 3 ) > $O/spgemm_gpu_$N.txt 2>&1; tail -2 $O/spgemm_gpu_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --workload small --steps 3 --warmup 3 --validate > $O/bench_small_$N.json 2> $O/bench_small_$N.err; echo "exit $?" >> $O/bench_small_$N.err; tail -2 $O/bench_small_$N.err
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_cfg3_$N.json 2> $O/bench_cfg3_$N.err; echo "exit $?" >> $O/bench_cfg3_$N.err; tail -2 $O/bench_cfg3_$N.err
cat $O/bench_cfg3_$N.json | cut -c1-600

