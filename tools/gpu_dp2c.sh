#!/bin/bash
# 2-GPU sanity of the final round-1 code: data-parallel bench (peer all-reduce + NCCL buckets + multi-tensor Adam)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_dp2c.log 2>&1
echo "dp2 exit $?"; tail -n 1 gpurun_out/bench_dp2c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','launch_mode','gpu_launches','loss')}, d['e2e']['value'] if d['e2e'] else None)"
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_dp2c.log 2>&1
echo "ref exit $?"; tail -n 1 gpurun_out/bench_ref_dp2c.log | cut -c1-200
