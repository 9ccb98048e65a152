#!/bin/bash
# A/B pass on one GPU:  gpurun --timeout 2400 -- 'bash tools/gpu_ab.sh'
#   kernel / conv / model tests at the current defaults (and with the packed stem-pool forward), then bench lines with the
#   switches under test.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_kernels.py tests/test_gpu_dropout.py -m gpu -x -q -p no:cacheprovider > gpurun_out/tests_k.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/tests_k.log
ADNI_POOL_STREAM=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "stem_tail or maxpool" -p no:cacheprovider > gpurun_out/tests_pool2.log 2>&1; echo "pool2 tests exit $?"; tail -n 3 gpurun_out/tests_pool2.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/tests_models.log 2>&1; echo "model tests exit $?"; tail -n 3 gpurun_out/tests_models.log
b() { tag=$1; shift; timeout 400 python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag exit $?"; python - gpurun_out/bench_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d['roofline']['kernels']; h=d['roofline']['hbm_kernels']
    print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms', {n:round(v['kernel_ms_per_step'],2) for n,v in k.items() if v['kernel_ms_per_step']>0}, {n:round(v['kernel_ms_per_step'],2) for n,v in h.items() if n.startswith('pool') or n.startswith('bn_bwd_red') or n.startswith('relu')})
except Exception as e: print('no line',e)
PY
}
b b32 --shape-profile gpurun_out/shapes_b32.json
ADNI_POOL_STREAM=2 b b32_pool2
b b32_again
b faithful --workload pet_mri_fusion_faithful --steps 4 --shape-profile gpurun_out/shapes_faithful.json
b b4 --global-batch 4 --steps 20
