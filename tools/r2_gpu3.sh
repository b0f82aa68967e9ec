#!/bin/bash
# canary first: a hang must not eat the whole call
mkdir -p gpurun_out
timeout 90 python tools/r2_quick.py "" > gpurun_out/r2_quick3.log 2>&1; rc=$?
tail -4 gpurun_out/r2_quick3.log
if [ $rc -ne 0 ]; then echo "canary failed rc=$rc"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_glm.py tests/test_gpu_fused_step.py tests/test_gpu_full_size.py tests/test_gpu_api.py -m gpu -q -x --timeout 120 > gpurun_out/r2_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
grep -E "FAILED|passed|failed|rc=|Error" gpurun_out/r2_pytest3.log | head -20
PATHS=tc_parity timeout 90 python tools/r2_timeline.py > gpurun_out/r2_timeline2.log 2>&1; head -8 gpurun_out/r2_timeline2.log
timeout 200 python bench.py --steps 1000 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench3.json"))
print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "launches", d["gpu_launches"])
print("roofline", d["roofline"]["us_per_launch"], d["roofline"]["frac"])
print("tensor", d["roofline_tensor"]["us_per_call"], d["roofline_tensor"]["frac"])
print("e2e", d["e2e"]["value"], "clocks", d["clocks"])
PY
tail -3 gpurun_out/r2_bench3.err
