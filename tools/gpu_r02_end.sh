#!/bin/bash
# tools/gpu_r02_end.sh — last GPU call of the round: the full GPU suite and the smoke with the final library.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02end_gputests.log 2>&1; echo "gpu tests exit $?"; tail -3 $O/r02end_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
