#!/bin/bash
# Round-2 GPU session D: new rows (mass matrix adaption, callable probe, tempering options)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mh_samplers.py tests/test_gpu_api.py tests/test_gpu_tempering.py tests/test_gpu_updates.py -m gpu -q --timeout 200 2>&1 | tail -40
