#!/bin/bash
# round-2 GPU call 2: CTA-pair potential kernel (cta_group::2) -- correctness, then timings
mkdir -p gpurun_out
timeout 120 python tools/r2_quick.py "5=1" > gpurun_out/r2_quick_cg1.log 2>&1; echo "cg1 rc=$?" >> gpurun_out/r2_quick_cg1.log
cat gpurun_out/r2_quick_cg1.log
timeout 120 python tools/r2_quick.py "" > gpurun_out/r2_quick_cg2.log 2>&1; echo "cg2 rc=$?" >> gpurun_out/r2_quick_cg2.log
cat gpurun_out/r2_quick_cg2.log
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv,noheader
timeout 420 python -m pytest tests/test_gpu_glm.py tests/test_gpu_fused_step.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -15 gpurun_out/r2_pytest2.log
for o in "" "5=1" "4=1"; do
  echo "== SGMC_OPTIONS=$o" >> gpurun_out/r2_glm2.log
  SGMC_OPTIONS=$o timeout 120 python tools/bench_glm.py --paths tc_parity,tc_throughput --observations 1000000 >> gpurun_out/r2_glm2.log 2>&1
done
cat gpurun_out/r2_glm2.log
timeout 300 python bench.py --steps 500 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
cat gpurun_out/r2_bench2.json; tail -3 gpurun_out/r2_bench2.err
