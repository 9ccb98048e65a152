#!/bin/bash
# Evidence for profiles/: launch list of the bench command and DRAM traffic + full sections of the dominant kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu launch list exit $?"; wc -l gpurun_out/launches.csv
CMD2="python tools/profile_kernels.py 32"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"igemm_kmajor|igemm_halo|wgrad2_kernel|wgrad_halo_kernel|stem_fprop_plane|stem_wgrad_plane" -s 14 -c 14 -o gpurun_out/prof_conv $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2>/dev/null
SZ=$(stat -c %s gpurun_out/prof_conv.ncu-rep); if [ "$SZ" -gt 40000000 ]; then rm gpurun_out/prof_conv.ncu-rep; echo "rep too large ($SZ), kept csv only"; fi
ls -la gpurun_out | head -20
