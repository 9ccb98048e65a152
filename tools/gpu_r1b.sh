#!/bin/bash
# round-1 session-3 check: new Adam kernel + pooled wgrad accumulators + M64 stem wgrad: tests, then a short bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/tests_r1b.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/tests_r1b.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1b.log 2>&1
echo "bench exit $?"; tail -n 1 gpurun_out/bench_r1b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','launch_mode','gpu_launches')}, d['e2e']['value'], d['eager'], d['roofline']['frac'])"
grep -v '^{' gpurun_out/bench_r1b.log | tail -5
