#!/bin/bash
# A/B pass on one GPU:  gpurun --timeout 2400 -- 'bash tools/gpu_ab.sh'
#   conv / kernel / model tests at the current defaults, then the default bench and the 8-GPU-per-rank-batch proxy
#   (--global-batch 4) with the stream-K schedules on / off (ADNI_STREAM_K = fprop+dgrad, ADNI_STREAM_K_WGRAD = wgrad).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_kernels.py -m gpu -x -q -p no:cacheprovider > gpurun_out/tests_sk.log 2>&1; echo "tests exit $?"; tail -n 5 gpurun_out/tests_sk.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/tests_models_sk.log 2>&1; echo "model tests exit $?"; tail -n 3 gpurun_out/tests_models_sk.log
b() { tag=$1; shift; timeout 400 python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag exit $?"; python - gpurun_out/bench_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d['roofline']['kernels']
    print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms', {n:round(v['kernel_ms_per_step'],2) for n,v in k.items() if v['kernel_ms_per_step']>0})
except Exception as e: print('no line',e)
PY
}
b b32_all --shape-profile gpurun_out/shapes_b32_sk.json
ADNI_STREAM_K=0 b b32_wgradsk
ADNI_STREAM_K=0 ADNI_STREAM_K_WGRAD=0 b b32_static
b b4_all --global-batch 4 --steps 20 --shape-profile gpurun_out/shapes_b4_sk.json
ADNI_STREAM_K=0 b b4_wgradsk --global-batch 4 --steps 20
ADNI_STREAM_K=0 ADNI_STREAM_K_WGRAD=0 b b4_static --global-batch 4 --steps 20
b faithful --workload pet_mri_fusion_faithful --steps 4 --shape-profile gpurun_out/shapes_faithful.json
b r50 --workload mri_r50_160 --steps 3 --warmup 2
