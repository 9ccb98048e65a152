#!/bin/bash
# session-3 check 2: full suite without the tiny-map conv cases, then the tiny cases in their own process
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "not tiny" > gpurun_out/tests_r1c.log 2>&1
echo "tests exit $?"; tail -n 12 gpurun_out/tests_r1c.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 -p no:cacheprovider -k "tiny" > gpurun_out/tests_tiny.log 2>&1
echo "tiny exit $?"; tail -n 8 gpurun_out/tests_tiny.log | cut -c1-300
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c.log 2>&1
echo "bench exit $?"; tail -n 1 gpurun_out/bench_r1c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','launch_mode','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
for k,v in d['roofline']['hbm_kernels'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})"
