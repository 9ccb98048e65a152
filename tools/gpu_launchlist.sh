#!/bin/bash
# ncu launch list (device time + DRAM bytes per launch) of one step:  gpurun -- 'bash tools/gpu_launchlist.sh <tag> <bench args>'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
tag=$1; shift
CMD="python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline $@"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launches_$tag.log 2>&1
echo "launch list $tag exit $?"
