#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/k.log 2>&1
echo "conv tests exit $?"; tail -n 2 gpurun_out/k.log
ADNI_DEBUG_MODE=0 timeout 300 python tools/fprop_probe.py child > gpurun_out/probe.txt 2>&1; cat gpurun_out/probe.txt
