#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
echo "bench full rc=$? ($(( $(date +%s) - t0 )) s)"; tail -n 3 gpurun_out/r02_bench_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-tebd --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "ncu rc=$?"
