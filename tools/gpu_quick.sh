#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conv.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/tests.log 2>&1
echo "tests exit $?"; tail -n 2 gpurun_out/tests.log
timeout 300 python tools/conv_shapes_probe.py > gpurun_out/conv_shapes.log 2>&1
echo "probe exit $?"; grep -v Warn gpurun_out/conv_shapes.log | tail -12
timeout 300 python tools/halo_probe.py > gpurun_out/halo.log 2>&1
echo "halo probe exit $?"; grep -v Warn gpurun_out/halo.log | grep shape | tail -6
timeout 1200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --shape-profile gpurun_out/shapes.json > gpurun_out/bench_full.log 2>&1
echo "bench full exit $?"; tail -n 1 gpurun_out/bench_full.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','launch_mode')}, d['eager']['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_executed'])"
python - <<'PY'
import json
d=json.load(open('gpurun_out/shapes.json'))['__entry_points__']
for k,v in sorted(d.items(), key=lambda kv:-kv[1]['ms_per_step'])[:16]: print(f"  {k:36s} {v['calls_per_step']:5.0f} calls {v['ms_per_step']:7.3f} ms")
PY
