#!/bin/bash
# tools/gpu_r02_end.sh — last GPU call of the round: the full GPU suite and the smoke with the final library, the published workload's bench line.
O=gpurun_out; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02end_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/r02end_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --workload pub_n5e6_d5 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02end_pub.json 2> $O/r02end_pub.err; echo "pub exit $?"
python -c "
import json; d=json.loads(open('gpurun_out/r02end_pub.json').read().strip().splitlines()[-1]); print('pub %.3f ms | %s | validated %s' % (d['ms_per_step'], d['roofline']['kernel'], d['validated']['ok']))"
