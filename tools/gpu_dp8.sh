#!/bin/bash
# 8-GPU pass:  gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_dp8.sh'
#   the default bench (configs[2], strong scaling, 4 pairs per GPU) with the gradient-exchange variants (NVLink kernel,
#   NCCL, side-stream overlap) and config 5 (ResNet-50, 160x192x160, 8 per GPU, sync-BN, weak scaling) on one 8 x B200 node.
N=8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms','e2e',d['e2e'] and round(d['e2e']['value'],1),d['gradient_exchange'],d['sync_bn_exchange'],'first loss',d['parity']['first_step_loss'],d['parity']['first_step_logits_checksum'],d['clocks'])
except Exception as e: print('no line',e)
PY
}
run 29551 --steps 30 --warmup 5 --shape-profile gpurun_out/shapes_dp8.json > gpurun_out/bench_dp8.json 2> gpurun_out/bench_dp8.err; echo "dp8 (NVLink gradient kernel) exit $?"; show gpurun_out/bench_dp8.json; tail -n 2 gpurun_out/bench_dp8.err | cut -c1-300
ADNI_PEER_GRADS=0 run 29552 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_dp8_nccl.json 2> gpurun_out/bench_dp8_nccl.err; echo "dp8 NCCL gradients exit $?"; show gpurun_out/bench_dp8_nccl.json
ADNI_OVERLAP_GRADS=1 run 29553 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_dp8_overlap.json 2> gpurun_out/bench_dp8_overlap.err; echo "dp8 overlap exit $?"; show gpurun_out/bench_dp8_overlap.json
ADNI_PEER_REDUCE=0 ADNI_PEER_GRADS=0 run 29554 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_dp8_allnccl.json 2> gpurun_out/bench_dp8_allnccl.err; echo "dp8 everything through NCCL exit $?"; show gpurun_out/bench_dp8_allnccl.json
run 29555 --workload mri_r50_160 --steps 5 --warmup 3 --shape-profile gpurun_out/shapes_r50_dp8.json > gpurun_out/bench_r50_dp8.json 2> gpurun_out/bench_r50_dp8.err; echo "r50 dp8 exit $?"; show gpurun_out/bench_r50_dp8.json; tail -n 2 gpurun_out/bench_r50_dp8.err | cut -c1-300
