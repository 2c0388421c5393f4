#!/bin/bash
# tools/gpu_r02_final.sh — end of round 2 on one GPU: smoke, the full GPU suite, the bench lines of every workload (default run first: config 3 with
# the CPU baseline and the reference arm), then the ncu launch list and one --set full capture of the shipped config-3 kernel.
O=gpurun_out; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02fin_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/r02fin_smoke.log
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r02fin_gputests.log 2>&1; echo "gpu tests exit $?"; tail -4 $O/r02fin_gputests.log
timeout 900 python bench.py > $O/r02fin_cfg3.json 2> $O/r02fin_cfg3.err; echo "bench default exit $?"
timeout 900 python bench.py --impl reference > $O/r02fin_reference_arm.json 2> $O/r02fin_reference_arm.err; echo "reference arm exit $?"
for W in cfg2 cfg5 pub_n5e6_d5; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r02fin_$W.json 2> $O/r02fin_$W.err; echo "$W exit $?"
done
timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --validate-rows 6000 > $O/r02fin_cfg4.json 2> $O/r02fin_cfg4.err; echo "cfg4 exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02fin_launches.csv $CMD > $O/r02fin_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_sort_async" -s 12 -c 1 -f -o $O/r02fin_cfg3_kernel $CMD > $O/r02fin_ncu_full.log 2>&1; echo "ncu full exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02fin_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl") == "reference":
            print(f.split("/")[-1], "reference arm: %.3e %s" % (d["value"], d["unit"]), d["cpu_baseline"]["sample"][:80]); continue
        e = d.get("e2e") or {}
        print(f.split("/")[-1], "ms/step %.4f kernel %.4f frac %.3f | validated %s | e2e %s | cpu %s" % (
            d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], (d["validated"] or {}).get("ok"), ("%.1f ms" % e["ms_per_step"]) if e else None,
            ("%.3e" % d["cpu_baseline"]["value"]) if d.get("cpu_baseline") else None))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
