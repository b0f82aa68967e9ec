#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_fused_step.py tests/test_gpu_api.py -m gpu -q -x --timeout 60 > gpurun_out/r2_pytest9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log
grep -E "FAILED|passed|failed|rc=|Error|Timeout" gpurun_out/r2_pytest9.log | head -10
if ! grep -q "rc=0" gpurun_out/r2_pytest9.log; then tail -30 gpurun_out/r2_pytest9.log; fi
for o in "" "9=1" "8=1" "8=1,9=1" "7=128"; do
  SGMC_OPTIONS=$o timeout 120 python tools/bench_scan.py --steps 2000 --reps 2 2>&1 | tail -2
done
timeout 200 python bench.py --steps 1000 --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench9.json"))
print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "launches", d["gpu_launches"], "e2e", d["e2e"]["value"], d["clocks"])
PY
