#!/bin/bash
# session-3 final evidence: full GPU suite, a complete bench line, the ncu launch list (time + DRAM bytes) of one step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/tests_r1d.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/tests_r1d.log | cut -c1-300
timeout 400 python bench.py --steps 6 --warmup 3 --shape-profile gpurun_out/shapes_r1d.json > gpurun_out/bench_r1d.log 2>&1
echo "bench exit $?"; tail -n 1 gpurun_out/bench_r1d.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus','launch_mode','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])
for k,v in d['roofline']['hbm_kernels'].items(): print(k, round(v['frac'],3), round(v['kernel_ms_per_step'],3))"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-graph"
$CMD > gpurun_out/plain_r1d.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 1100 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu_r1d.log 2>&1
echo "ncu launch list exit $?"; wc -l gpurun_out/launches_r1d.csv
