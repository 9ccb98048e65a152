#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?"; tail -n 2 gpurun_out/tests.log
timeout 300 python tools/conv_shapes_probe.py 2>&1 | grep "^N32" | head -4
for i in 1 2 3 4 5 6; do
timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/repro_$i.log 2>&1
rc=$?
echo "run $i exit $rc $(tail -n 1 gpurun_out/repro_$i.log | cut -c60-150)"
grep "adni_b200" gpurun_out/repro_$i.log | head -4
if [ $rc -ne 0 ]; then break; fi
done
