#!/bin/bash
mkdir -p gpurun_out
bash tools/r2_gpu6.sh 2>&1 | tail -6 | cut -c1-220
timeout 120 python tools/r2_cold_start_amplification.py > gpurun_out/r2_cold_start.log 2>&1; cat gpurun_out/r2_cold_start.log
timeout 400 python -m pytest tests/test_gpu_full_size.py -m gpu -q --timeout 200 -k "trajectory or determinism" > gpurun_out/r2_pytest8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
grep -E "FAILED|passed|failed|rc=|Error|Timeout|assert" gpurun_out/r2_pytest8.log | head -20
