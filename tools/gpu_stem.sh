#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "stem" > gpurun_out/tests_stem.log 2>&1
echo "tests exit $?"; tail -n 2 gpurun_out/tests_stem.log
timeout 300 python tools/stem_probe.py 2>&1 | grep -v Warn | tail -12
