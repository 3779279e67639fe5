#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_headline_parity.py -q -m gpu -k "c4 or c3 or c5 or eigh" --durations=6 > gpurun_out/r02_hp.log 2>&1
echo "rc=$?"; tail -n 30 gpurun_out/r02_hp.log
