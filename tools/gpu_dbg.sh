#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 200 python bench.py --steps 3 --warmup 3 --no-sweep --no-tebd --no-cpu-baseline > gpurun_out/r02_bench_q3.json 2> gpurun_out/r02_bench_q3.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02_bench_q3.err
python -c "
import json; j=json.loads(open('gpurun_out/r02_bench_q3.json').read().strip().splitlines()[-1]); print(j['value'], j['roofline']['frac'], j['e2e']['value'], j['parity']['ok'], j['roofline']['kernel_family_calls'], j['roofline']['plan_cache'])"
