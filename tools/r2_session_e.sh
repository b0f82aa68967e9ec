#!/bin/bash
# Round-2 GPU session E: apply-kernel changes (tests + timing), warm-cache DRAM traffic of the
# step's kernels, --set full capture of the stand-alone fused update (roofline.traffic)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_step.py tests/test_gpu_glm.py tests/test_gpu_updates.py tests/test_gpu_full_size.py -m gpu -q -x --timeout 300 2>&1 | tail -5
timeout 120 python tools/bench_scan.py --steps 2000 --reps 2 2>&1 | tail -1
timeout 120 python tools/r2_step_profile.py 2>&1 | tail -4
# warm caches: two metrics need one pass, nothing is flushed or replayed
timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:'k_glm_tc_pair|k_sgld_apply|k_prepare_all' --launch-skip 300 -c 30 --csv --log-file gpurun_out/r02_warm_traffic_raw.csv \
  python tools/bench_scan.py --steps 150 --reps 1 > /dev/null 2>&1; echo "ncu warm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_noise_pass' --launch-skip 8 -c 2 \
  -o gpurun_out/r02_update_standalone -f python tools/bench_update.py --kernels sgld_rms --reps 2 > gpurun_out/r02_update_standalone.log 2>&1; echo "ncu update rc=$?"
ls -la gpurun_out | tail -5
