#!/bin/bash
# Validation + A/B of the opt-in packed-compare stem pooling forward (csrc/stem_fused.cu, ADNI_POOL_STREAM=2):
#   gpurun --timeout 600 -- 'bash tools/gpu_pool_packed.sh'
# The existing tests assert bit-identical pooled values and arg-max bytes against the unfused kernels, so running
# them with the switch set validates the variant; then the bench with and without it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ADNI_POOL_STREAM=2 timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -x -q \
  -k "fused_stem or anat-0 or anat-1 or anat_pet_2resnet or config1" > gpurun_out/tests_pool_packed.log 2>&1
rc=$?
echo "packed pool tests exit $rc"; tail -n 4 gpurun_out/tests_pool_packed.log
if [ $rc -eq 0 ]; then
  for v in 1 2; do
    ADNI_POOL_STREAM=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pool$v.log 2>&1
    echo "ADNI_POOL_STREAM=$v"; tail -n 1 gpurun_out/bench_pool$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['hbm_kernels'].get('pool_fwd'))"
  done
fi
