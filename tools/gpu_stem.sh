#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -k tc_stem -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/stem.log 2>&1
echo "stem exit $?"; grep -E "^E  |passed|failed" gpurun_out/stem.log | head -40
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/models.log 2>&1
echo "models exit $?"; tail -n 4 gpurun_out/models.log
timeout 1200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
echo "bench full exit $?"; tail -n 2 gpurun_out/bench_full.log
