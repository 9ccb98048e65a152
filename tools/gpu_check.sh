#!/bin/bash
# One-GPU validation pass:  gpurun --timeout 2400 -- 'bash tools/gpu_check.sh [quick]'
#   GPU test suite, smoke(), the default bench line, then (unless "quick") the other BASELINE.json workloads, the
#   8-GPU-per-rank-batch proxy (--global-batch 4) with its per-shape table, and training from .nii.gz files.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/tests_all.log 2>&1; echo "gpu tests exit $?"; tail -n 4 gpurun_out/tests_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_default.json; tail -n 3 gpurun_out/bench_default.err
[ "$1" = "quick" ] && exit 0
timeout 900 python bench.py --workload mri_r50_160 --steps 3 --warmup 2 --shape-profile gpurun_out/shapes_r50.json > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "r50 exit $?"; cut -c1-400 gpurun_out/bench_r50.json; tail -n 5 gpurun_out/bench_r50.err
timeout 400 python bench.py --global-batch 4 --steps 10 --warmup 3 --no-cpu-baseline --shape-profile gpurun_out/shapes_b4.json > gpurun_out/bench_b4.json 2> gpurun_out/bench_b4.err; echo "b4 exit $?"; cut -c1-300 gpurun_out/bench_b4.json
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --shape-profile gpurun_out/shapes_b32.json > gpurun_out/bench_b32_shapes.json 2> gpurun_out/bench_b32_shapes.err; echo "b32 shapes exit $?"
for w in mri_r18 mri_r10 pet_mri_fusion_faithful all_modalities; do
  timeout 500 python bench.py --workload $w --steps 4 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"; cut -c1-300 gpurun_out/bench_$w.json; tail -n 2 gpurun_out/bench_$w.err
done
timeout 500 python tools/bench_staged_e2e.py > gpurun_out/staged_e2e.json 2> gpurun_out/staged_e2e.err; echo "staged e2e exit $?"; tail -c 1500 gpurun_out/staged_e2e.json
