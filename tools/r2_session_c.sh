#!/bin/bash
# Round-2 GPU session C: pull-mode tests, host-link experiment, bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_fused_step.py -m gpu -q -x --timeout 120 2>&1 | tail -15
timeout 200 python tools/exp_pull.py 2>&1 | tail -14
timeout 300 python bench.py --no-cpu-baseline --no-resgld > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
tail -3 gpurun_out/r02_bench_c.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c.json"))
print("value", d["value"], "us/step", d["ms_per_step"] * 1e3, "launches", d["gpu_launches"])
print("e2e", d["e2e"])
print(d["clocks"])
PY
SGMC_HOST_PULL=0 timeout 300 python bench.py --no-cpu-baseline --no-resgld 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('staged e2e', d['e2e']['value'], d['e2e']['host_link'])"
