#!/bin/bash
# One gpurun call: GPU tests (multi-rank suite + headline-size parity first), then a short bench.  Logs -> gpurun_out/.
# usage: tools/gpu_ci.sh [new] [rest] [bench] [benchfull]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
want() { [[ " $ARGS " == *" $1 "* ]]; }
ARGS="${*:-new rest bench}"
if want new; then
  t0=$(date +%s)
  timeout 900 python -m pytest tests/test_gpu_shard_dmrg.py tests/test_gpu_headline_parity.py -x -q -m gpu --durations=8 > gpurun_out/r02_tests_new.log 2>&1
  echo "new tests rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 25 gpurun_out/r02_tests_new.log
fi
if want rest; then
  t0=$(date +%s)
  timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_shard_dmrg.py --deselect tests/test_gpu_headline_parity.py > gpurun_out/r02_tests_rest.log 2>&1
  echo "rest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 5 gpurun_out/r02_tests_rest.log
fi
if want bench; then
  t0=$(date +%s)
  timeout 600 python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err
  echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"; tail -c 3000 gpurun_out/r02_bench_quick.json; tail -n 5 gpurun_out/r02_bench_quick.err
fi
if want benchfull; then
  t0=$(date +%s)
  timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
  echo "bench full rc=$? ($(( $(date +%s) - t0 )) s)"; tail -c 6000 gpurun_out/r02_bench_full.json; tail -n 5 gpurun_out/r02_bench_full.err
fi
