#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/c5_probe.py 2048 gpurun_out/r02_c5_chi2048.json > gpurun_out/c5_2048.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/c5_2048.log
timeout 400 python tools/c5_probe.py 8192 gpurun_out/r02_c5_chi8192.json > gpurun_out/c5_8192.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/c5_8192.log
