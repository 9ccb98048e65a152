#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do
timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/repro_$i.log 2>&1
rc=$?
echo "run $i exit $rc $(grep -c 'adni_b200' gpurun_out/repro_$i.log)"
grep "adni_b200" gpurun_out/repro_$i.log | head -4
if [ $rc -ne 0 ]; then break; fi
done
