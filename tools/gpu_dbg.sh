#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_heff.py -q -m gpu -k host > gpurun_out/r02_host.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r02_host.log
for nc in 4 8 16; do
TNB_HOST_CHUNKS=$nc timeout 200 python bench.py --steps 5 --warmup 3 --no-sweep --no-tebd --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $nc: value %.2f e2e %.2f (%.2f ms) host_path_rel_err %s'%(j['value'], j['e2e']['value'], j['e2e']['ms_per_step'], j['parity']['host_path_rel_err']))"
done
