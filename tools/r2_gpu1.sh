#!/bin/bash
# round-2 GPU call 1: correctness of the persistent fused potential kernel + first timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
for o in "" "5=256" "4=1"; do
  echo "== SGMC_OPTIONS=$o" >> gpurun_out/r2_glm1.log
  SGMC_OPTIONS=$o timeout 120 python tools/bench_glm.py --paths tc_parity,tc_throughput --observations 1000000 >> gpurun_out/r2_glm1.log 2>&1
done
cat gpurun_out/r2_glm1.log
timeout 300 python tools/r2_traj_err.py > gpurun_out/r2_traj.log 2>&1
cat gpurun_out/r2_traj.log
SGMC_OPTIONS=4=1 timeout 300 python tools/r2_traj_err.py > gpurun_out/r2_traj_legacy.log 2>&1
grep "tc_parity" gpurun_out/r2_traj_legacy.log
timeout 300 python bench.py --steps 500 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
cat gpurun_out/r2_bench1.json
