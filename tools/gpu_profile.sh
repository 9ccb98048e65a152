#!/bin/bash
# ncu evidence pass (one GPU):  gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh'
#   1. launch list (device time + DRAM bytes per launch) of one training step of the default bench workload
#   2. ncu --set full of the dominant tensor kernels (tap-per-box fprop/dgrad, wgrad2) in that workload
#   3. the same for the small-channel engine in the reference-faithful workload
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
CMD2="python bench.py --workload pet_mri_fusion_faithful --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"small_fprop_kernel|small_wgrad_kernel" -s 11 -c 11 \
    -o gpurun_out/prof_small -f $CMD2 > gpurun_out/ncu_small.log 2>&1
echo "small full exit $?"
ls -la gpurun_out/*.ncu-rep
