#!/bin/bash
# usage: gpu_ncu_one.sh <python script> <kernel regex> <output stem>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python $1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$2" -s 2 -c 1 -o gpurun_out/$3 $CMD > gpurun_out/ncu_one.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/$3.ncu-rep --page source --csv > gpurun_out/$3_source.csv 2>/dev/null
ncu -i gpurun_out/$3.ncu-rep --page raw --csv > gpurun_out/$3_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep $3
