cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
CMD="python bench.py --workload mri_r50_160 --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_r50.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"igemm_kmajor_kernel" -s 0 -c 14 -o gpurun_out/prof_r50_k1 -f $CMD > gpurun_out/ncu_r50.log 2>&1
echo "ncu r50 exit $?"; ls -la gpurun_out/*.ncu-rep
