#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_step.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for o in "" "12=1"; do
  SGMC_OPTIONS=$o timeout 120 python tools/bench_scan.py --steps 2000 --reps 2 2>&1 | tail -1
  timeout 120 python tools/r2_step_profile.py "$o" 2>&1 | tail -1
done
