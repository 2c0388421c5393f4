#!/bin/bash
# tools/gpu_r02_x.sh — full GPU suite, then compute-sanitizer (memcheck / racecheck / synccheck) on the small parity cases.
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r02x_gputests.log 2>&1; echo "gpu tests exit $?"; tail -4 $O/r02x_gputests.log
SEL="fixture_golden or kats or seeded_cases_golden or edge_cases or every_bin or big_rows_dense or masked_product"
for TOOL in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/r02x_sanitizer_$TOOL.log 2>&1; echo "$TOOL exit $?"; tail -3 $O/r02x_sanitizer_$TOOL.log
done
# racecheck: shared-memory hazards.  rows_bm.cuh's pass A races on purpose (plain read-modify-write, repaired by pass B), so the big-row cases are left out here.
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixture_golden or kats or seeded_cases_golden or edge_cases or masked_product" > $O/r02x_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 $O/r02x_sanitizer_racecheck.log
