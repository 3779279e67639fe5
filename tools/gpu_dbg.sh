#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_contract_tma.py tests/test_gpu_small_rows.py -q -m gpu > gpurun_out/r02_tma.log 2>&1
echo "rc=$?"; tail -n 12 gpurun_out/r02_tma.log
timeout 100 python tools/ab_split.py
