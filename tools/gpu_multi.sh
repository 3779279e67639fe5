#!/bin/bash
# usage: tools/gpu_multi.sh N [c5] [extra bench args]   -- multi-GPU validation: NCCL-group suite for world=N, bench --gpus N, optional C5 probe
N=$1; shift
C5=0; if [ "$1" == "c5" ]; then C5=1; shift; fi
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
t0=$(date +%s)
timeout 400 python -m pytest "tests/test_gpu_shard_dmrg.py::test_multi_rank_suite[$N]" -x -q -m gpu > gpurun_out/r02_suite_n$N.log 2>&1
echo "suite[$N] rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 4 gpurun_out/r02_suite_n$N.log
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench N=$N rc=$? ($(( $(date +%s) - t0 )) s)"; tail -c 5000 gpurun_out/r02_bench_n$N.json; grep -v "^W\|^\*\*\*\|OMP_NUM\|unbatched P2P" gpurun_out/r02_bench_n$N.err | tail -n 8
if [ $C5 == 1 ]; then
  t0=$(date +%s)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/c5_probe_multi.py 8192 gpurun_out/r02_c5_chi8192_n$N.json > gpurun_out/c5_multi.log 2>&1
  echo "c5 multi rc=$? ($(( $(date +%s) - t0 )) s)"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/c5_multi.log | tail -n 25
fi
