#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/r2_quick.py "" > gpurun_out/r2_quick7.log 2>&1; rc=$?
tail -3 gpurun_out/r2_quick7.log
if [ $rc -ne 0 ]; then echo "canary failed rc=$rc"; exit 1; fi
timeout 200 python -m pytest tests/test_gpu_fused_step.py -m gpu -q -x --timeout 60 > gpurun_out/r2_pytest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
grep -E "FAILED|passed|failed|rc=|Error|Timeout" gpurun_out/r2_pytest7.log | head -20
for o in "" "8=1"; do
  echo "=== SGMC_OPTIONS=$o"
  SGMC_OPTIONS=$o MODE=step PATHS=tc_parity timeout 90 python tools/r2_timeline.py > gpurun_out/r2_timeline7_$o.log 2>&1; head -5 gpurun_out/r2_timeline7_$o.log | cut -c1-200
  SGMC_OPTIONS=$o timeout 200 python bench.py --steps 1000 --no-cpu-baseline > gpurun_out/r2_bench7_$o.json 2> gpurun_out/r2_bench7.err
  python - "$o" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2_bench7_{sys.argv[1]}.json"))
print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "launches", d["gpu_launches"], "tensor", d["roofline_tensor"]["us_per_call"], "e2e", d["e2e"]["value"], d["clocks"])
PY
done
timeout 300 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2_pytest7b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest7b.log
grep -E "FAILED|passed|failed|rc=|Error|Timeout" gpurun_out/r2_pytest7b.log | head -20
