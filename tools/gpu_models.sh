#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/models.log 2>&1
echo "models exit $?"
tail -n 60 gpurun_out/models.log
