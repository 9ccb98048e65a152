#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/tests_all.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/tests_all.log
timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --shape-profile gpurun_out/shapes.json > gpurun_out/bench_full.log 2>&1
echo "bench full exit $?"; tail -n 1 gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','launch_mode','gpu_launches')}, d['eager']['ms_per_step'])
r=d['roofline']; print({k:r[k] for k in ('achieved','executed','frac','frac_executed','peak')})
for k,v in r['other_kernels'].items(): print(' ',k[:40],v)"
python - <<'PY'
import json
d=json.load(open('gpurun_out/shapes.json'))['__entry_points__']
for k,v in sorted(d.items(), key=lambda kv:-kv[1]['ms_per_step'])[:18]: print(f"  {k:36s} {v['calls_per_step']:5.0f} calls {v['ms_per_step']:7.3f} ms")
PY
