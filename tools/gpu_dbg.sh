#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_contract_tma.py -x -q -m gpu > gpurun_out/r02_tma.log 2>&1
echo "rc=$?"; tail -n 40 gpurun_out/r02_tma.log
