#!/bin/bash
# Round-2 GPU session B: pair-kernel phase timeline (carried step) for the tile widths,
# scan timing per option, new tempering tests.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tempering.py tests/test_gpu_api.py -m gpu -q -x --timeout 100 2>&1 | tail -3
for o in "" "7=128"; do
  echo "### timeline options=$o"
  SGMC_OPTIONS=$o MODE=step PATHS=tc_parity timeout 120 python tools/r2_timeline.py 2>&1 | tee gpurun_out/r02_timeline_${o//=/_}.txt | head -50
done
for o in "" "7=128" "8=1" "9=1"; do
  SGMC_OPTIONS=$o timeout 120 python tools/bench_scan.py --steps 2000 --reps 2 2>&1 | tail -1
done
