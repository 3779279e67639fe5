#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_small_rows.py tests/test_gpu_contract.py tests/test_gpu_vecops.py -q -m gpu --durations=4 > gpurun_out/r02_small.log 2>&1
echo "rc=$?"; tail -n 60 gpurun_out/r02_small.log
