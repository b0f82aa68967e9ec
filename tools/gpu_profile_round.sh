#!/bin/bash
# One GPU session: parity tests, the bench line, the ncu launch list of the same
# bench command and one --set full capture of the step's kernels.
# Usage (from the repo root, on the GPU box): bash tools/gpu_profile_round.sh rNN
R=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
cat gpurun_out/${R}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${R}_launches_raw.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${R}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_glm_tc_gemm|k_noise_pass|k_prepare_all|k_minibatch' --launch-skip 40 -c 10 \
  -o gpurun_out/${R}_step_kernels -f python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
