#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for m in 0 1; do
ADNI_STEM_M64=$m timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "stem" > gpurun_out/tests_stem_$m.log 2>&1
echo "M64=$m tests exit $?"; tail -n 2 gpurun_out/tests_stem_$m.log | head -1
ADNI_STEM_M64=$m timeout 300 python tools/stem_probe.py 2>&1 | grep "stem wgrad"
done
