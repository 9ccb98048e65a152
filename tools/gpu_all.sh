#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/tests.log
timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
echo "bench full exit $?"; tail -n 1 gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','launch_mode','gpu_launches')}, d['e2e']['value'], d['eager'], d['roofline']['frac'])"
