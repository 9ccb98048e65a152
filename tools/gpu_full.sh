#!/bin/bash
# full GPU test suite + the default bench line (e2e + cpu_baseline) + the reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/tests_all.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/tests_all.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
echo "bench exit $?"; tail -n 1 gpurun_out/bench_default.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "ref exit $?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
