#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mode in 0 1 1; do
ADNI_PEER_REDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dp2_$mode.log 2>&1
echo "peer=$mode exit $?"; grep "adni_b200" gpurun_out/bench_dp2_$mode.log | head -3; tail -n 1 gpurun_out/bench_dp2_$mode.log | cut -c1-160
done
