#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/profile_one.py"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"igemm_halo" -s 2 -c 1 -o gpurun_out/prof_halo $CMD > gpurun_out/ncu_one.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_halo.ncu-rep --page source --csv > gpurun_out/prof_halo_source.csv 2>/dev/null
ncu -i gpurun_out/prof_halo.ncu-rep --page raw --csv > gpurun_out/prof_halo_raw.csv 2>/dev/null
ls -la gpurun_out/ | head
