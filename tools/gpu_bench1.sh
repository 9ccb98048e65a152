#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/models.log 2>&1
echo "models exit $?"; tail -n 5 gpurun_out/models.log
timeout 900 python bench.py --steps 3 --warmup 2 --global-batch 8 --no-cpu-baseline > gpurun_out/bench_b8.log 2>&1
echo "bench b8 exit $?"; tail -n 5 gpurun_out/bench_b8.log
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_full.log 2>&1
echo "bench full exit $?"; tail -n 5 gpurun_out/bench_full.log
