#!/bin/bash
# ncu evidence pass (one GPU):  gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'
#   1. launch list (device time + DRAM bytes per launch) of one training step of the default bench workload
#   2. ncu --set full of the dominant tensor kernels (tap-per-box fprop/dgrad, wgrad2) in that workload
#   3. the same for the first tap-per-box launches of the ResNet-50 step (the flat 1x1x1 path)
# Every profiled command line first runs plain (B200_PROFILING.md); numbers printed under ncu are never bench values.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; grep -c igemm gpurun_out/launches.csv
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"igemm_kmajor_kernel|wgrad2_kernel" -s 60 -c 12 \
    -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_conv.log 2>&1
echo "conv full exit $?"
CMD3="python bench.py --workload mri_r50_160 --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline"
$CMD3 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"igemm_kmajor_kernel" -s 0 -c 10 \
    -o gpurun_out/prof_r50_k1 -f $CMD3 > gpurun_out/ncu_r50.log 2>&1
echo "r50 flat full exit $?"
ls -la gpurun_out/*.ncu-rep
