#!/bin/bash
# Round-2 GPU session A: smoke, full GPU parity suite, bench line, ncu launch
# list of the bench command, one --set full capture of the step's kernels.
R=r02
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "FAILED|ERROR|passed|failed" gpurun_out/${R}_pytest_gpu.log | head -30
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
cat gpurun_out/${R}_bench.json | cut -c1-3000
tail -5 gpurun_out/${R}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${R}_launches_raw.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${R}_bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_glm_tc|k_noise_pass|k_sgld_apply|k_prepare_all|k_randint' --launch-skip 40 -c 12 \
  -o gpurun_out/${R}_step_kernels -f python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${R}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
