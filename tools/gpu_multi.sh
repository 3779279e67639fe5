#!/bin/bash
# usage: tools/gpu_multi.sh N [extra bench args]   -- multi-GPU validation: NCCL-group suite for world=N, then bench --gpus N
N=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
t0=$(date +%s)
timeout 400 python -m pytest "tests/test_gpu_shard_dmrg.py::test_multi_rank_suite[$N]" -x -q -m gpu > gpurun_out/r02_suite_n$N.log 2>&1
echo "suite[$N] rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 4 gpurun_out/r02_suite_n$N.log
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench N=$N rc=$? ($(( $(date +%s) - t0 )) s)"; tail -c 5000 gpurun_out/r02_bench_n$N.json; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02_bench_n$N.err | tail -n 8
