#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused_step.py -m gpu -q -x --timeout 300 2>&1 | tail -4
for o in "" "13=1"; do
  for i in 1 2; do SGMC_OPTIONS=$o timeout 120 python tools/bench_scan.py --steps 3000 --reps 2 2>&1 | tail -1; done
done
SGMC_OPTIONS=13=1 ALL_PAIRS=0 MODE=step PATHS=tc_parity timeout 120 python tools/r2_timeline.py 2>&1 | grep -E 'slot  1:|slot 28|slot 29|slot 31'
