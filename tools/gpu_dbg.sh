#!/bin/bash
mkdir -p gpurun_out
export TNB_TEST_TRACE=$PWD/gpurun_out
timeout 300 python -m pytest tests/test_gpu_dmrg_lockstep.py -q -m gpu > gpurun_out/r02_lock.log 2>&1
echo "rc=$?"; tail -n 5 gpurun_out/r02_lock.log
