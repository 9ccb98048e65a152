#!/bin/bash
# round-1 final check, the driver's own sequence: full GPU suite, smoke(), default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tests_r1f.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/tests_r1f.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1f.log 2>&1
echo "smoke exit $?"; tail -n 4 gpurun_out/smoke_r1f.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/bench_r1f.log 2>&1
echo "bench exit $?"; tail -n 1 gpurun_out/bench_r1f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','steps','launch_mode','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])
for k,v in d['roofline']['hbm_kernels'].items(): print(k, round(v['frac'],3), round(v['kernel_ms_per_step'],3))"
