#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_parity.py > gpurun_out/dp_parity.log 2>&1
echo "dp parity exit $?"; grep -v Warn gpurun_out/dp_parity.log | tail -n 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --shape-profile gpurun_out/shapes_dp2.json > gpurun_out/bench_dp2.log 2>&1
echo "dp2 exit $?"; tail -n 1 gpurun_out/bench_dp2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','launch_mode','gpu_launches')}, d['e2e']['value'] if d['e2e'] else None)"
python - <<'PY'
import json
d=json.load(open('gpurun_out/shapes_dp2.json'))['__entry_points__']
for k,v in sorted(d.items(), key=lambda kv:-kv[1]['ms_per_step'])[:30]:
    if 'peer' in k or 'bn_' in k or 'conv3d' in k: print(f"  {k:36s} {v['calls_per_step']:5.0f} calls {v['ms_per_step']:7.3f} ms")
PY
