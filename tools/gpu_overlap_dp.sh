#!/bin/bash
# A/B of the opt-in overlapped gradient all-reduce (data_parallel.OverlappedGradientBuckets, ADNI_OVERLAP_GRADS=1) on N GPUs:
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_overlap_dp.sh 2'
# The loss printed by both runs must agree (same seeds, same reductions); compare value / ms_per_step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
for mode in 0 1; do
  ADNI_OVERLAP_GRADS=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $((29540 + mode)) bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_overlap${mode}_dp$N.log 2>&1
  echo "overlap=$mode exit $?"; tail -n 1 gpurun_out/bench_overlap${mode}_dp$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','launch_mode','loss')})"
done
