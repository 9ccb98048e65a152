#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp$N.log 2>&1
echo "dp$N exit $?"; tail -n 1 gpurun_out/bench_dp$N.log | cut -c1-200
