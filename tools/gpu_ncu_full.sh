#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/profile_kernels.py 8"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"igemm_kmajor|wgrad2|stem_fprop|stem_wgrad" -s 14 -c 14 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
ls -la gpurun_out/
SZ=$(stat -c %s gpurun_out/prof_conv.ncu-rep); if [ "$SZ" -gt 45000000 ]; then rm gpurun_out/prof_conv.ncu-rep; echo "rep too large ($SZ), kept csv only"; fi
