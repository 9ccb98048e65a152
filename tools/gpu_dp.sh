#!/bin/bash
# N-GPU pass:  gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_dp.sh N [extra] [notests]'
#   oracle-based multi-rank parity tests (2 ranks; they include the NVLink exchange kernels against NCCL), the default
#   bench at N GPUs, the 8-GPU per-rank batch (4 pairs per rank) on N ranks with the exchange variants, and (with
#   "extra") config 5 (ResNet-50, 160x192x160, 8 per GPU, sync-BN) at N GPUs.
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms','e2e',d['e2e'] and round(d['e2e']['value'],1),d['gradient_exchange'],'|',d['sync_bn_exchange'],'| first loss',d['parity']['first_step_loss'],d['parity']['first_step_logits_checksum'])
except Exception as e: print('no line',e)
PY
}
[ "$3" = "notests" ] || { timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/tests_multirank.log 2>&1; echo "multirank tests exit $?"; grep -E "DP ORACLE|loss sharded|logits rel|gradients:|running stat|passed|failed|peer gradient|peer all-reduce|gradient slots|Error|error|timeout" gpurun_out/tests_multirank.log | tail -n 16; }
run 29541 --steps 20 --warmup 5 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "dp$N exit $?"; show gpurun_out/bench_dp$N.json
run 29544 --global-batch $((4*N)) --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_dp${N}_b4.json 2> gpurun_out/bench_dp${N}_b4.err; echo "dp$N b4/rank exit $?"; show gpurun_out/bench_dp${N}_b4.json
ADNI_OVERLAP_GRADS=0 run 29546 --global-batch $((4*N)) --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_dp${N}_b4_nooverlap.json 2> gpurun_out/bench_dp${N}_b4_nooverlap.err; echo "dp$N b4/rank, exchange after backward exit $?"; show gpurun_out/bench_dp${N}_b4_nooverlap.json
ADNI_PEER_GRADS=0 ADNI_OVERLAP_GRADS=0 run 29545 --global-batch $((4*N)) --steps 30 --warmup 5 --no-e2e > gpurun_out/bench_dp${N}_b4_nccl.json 2> gpurun_out/bench_dp${N}_b4_nccl.err; echo "dp$N b4/rank NCCL gradients exit $?"; show gpurun_out/bench_dp${N}_b4_nccl.json
if [ "$2" = "extra" ]; then
  run 29543 --workload mri_r50_160 --steps 5 --warmup 3 > gpurun_out/bench_r50_dp$N.json 2> gpurun_out/bench_r50_dp$N.err; echo "r50 dp$N exit $?"; show gpurun_out/bench_r50_dp$N.json
fi
