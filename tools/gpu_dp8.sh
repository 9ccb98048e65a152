#!/bin/bash
# 8-GPU pass:  gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_dp8.sh'
#   the default bench (configs[2], strong scaling, 4 pairs per GPU) and config 5 (ResNet-50, 160x192x160, 8 per GPU,
#   sync-BN, weak scaling) on one 8 x B200 node.
N=8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29551 --steps 30 --warmup 5 --shape-profile gpurun_out/shapes_dp8.json > gpurun_out/bench_dp8.json 2> gpurun_out/bench_dp8.err; echo "dp8 exit $?"; grep '^{' gpurun_out/bench_dp8.json | cut -c1-300; tail -n 2 gpurun_out/bench_dp8.err
run 29552 --workload mri_r50_160 --steps 5 --warmup 3 > gpurun_out/bench_r50_dp8.json 2> gpurun_out/bench_r50_dp8.err; echo "r50 dp8 exit $?"; grep '^{' gpurun_out/bench_r50_dp8.json | cut -c1-300; tail -n 2 gpurun_out/bench_r50_dp8.err
