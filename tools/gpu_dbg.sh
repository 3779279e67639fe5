#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest "tests/test_gpu_shard_dmrg.py::test_multi_rank_suite[2]" "tests/test_gpu_shard_dmrg.py::test_multi_rank_suite[4]" -x -q -m gpu > gpurun_out/r02_suite24.log 2>&1
echo "rc=$?"; tail -n 12 gpurun_out/r02_suite24.log
