#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for mode in 1 1 1 1 0; do
ADNI_PEER_REDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_dp2_$mode.log 2>&1
rc=$?
echo "peer=$mode exit $rc $(tail -n 1 gpurun_out/bench_dp2_$mode.log | cut -c60-120)"; grep "adni_b200" gpurun_out/bench_dp2_$mode.log | head -3
if [ $rc -ne 0 ]; then cp gpurun_out/bench_dp2_$mode.log gpurun_out/bench_dp2_fail.log; fi
done
