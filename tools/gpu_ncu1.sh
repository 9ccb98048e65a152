#!/bin/bash
# launch list of one short bench run (kernel time shares), after the same command ran clean without ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --global-batch 8 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
tail -n 3 gpurun_out/ncu.log
wc -l gpurun_out/launches.csv
