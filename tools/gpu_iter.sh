#!/bin/bash
# Iteration pass on one GPU:  gpurun --timeout 1500 -- 'bash tools/gpu_iter.sh'
#   conv / kernel tests, then bench lines (default, ResNet-50 config 5, the 8-GPU per-rank proxy) with per-shape tables.
#   Extra arguments are evaluated as one more command at the end.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_kernels.py -m gpu -x -q -p no:cacheprovider > gpurun_out/tests_k.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/tests_k.log
b() { tag=$1; shift; timeout 500 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag exit $?"; python - gpurun_out/bench_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); k=d['roofline']['kernels']; h=d['roofline']['hbm_kernels']
    print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms', 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['clocks'], {n:round(v['kernel_ms_per_step'],2) for n,v in k.items() if v['kernel_ms_per_step']>0}, {n:round(v['kernel_ms_per_step'],2) for n,v in h.items()})
except Exception as e: print('no line',e)
PY
}
python tools/k1_probe.py child > gpurun_out/k1_flat.txt 2>&1; cat gpurun_out/k1_flat.txt
ADNI_FLAT_1X1=0 b r50_box --no-e2e --workload mri_r50_160 --steps 4 --warmup 3
b r50 --workload mri_r50_160 --steps 4 --warmup 3 --shape-profile gpurun_out/shapes_r50.json
b b32 --steps 12 --shape-profile gpurun_out/shapes_b32.json
b b4 --no-e2e --global-batch 4 --steps 30
grep -h "launch_mode" gpurun_out/bench_b4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('launch mode', d['launch_mode'], 'parity', d['parity'])"
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests_all.log 2>&1; echo "all gpu tests exit $?"; tail -n 3 gpurun_out/tests_all.log
[ -n "$1" ] && eval "$@"
exit 0
