#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_small_rows.py -q -m gpu -k native > gpurun_out/r02_native.log 2>&1
echo "rc=$?"; tail -n 25 gpurun_out/r02_native.log
