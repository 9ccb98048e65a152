#!/bin/bash
# N-GPU pass:  gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_dp.sh N [extra]'
#   oracle-based multi-rank parity tests (2 ranks), the default bench at N GPUs with the plain and the overlapped
#   gradient exchange, and (with "extra") config 5 (ResNet-50, 160x192x160, 8 per GPU, sync-BN) at N GPUs.
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
[ "$3" = "notests" ] || timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s -p no:cacheprovider > gpurun_out/tests_multirank.log 2>&1; echo "multirank tests exit $?"; grep -E "DP ORACLE|loss sharded|logits rel|gradients:|running stat|passed|failed|peer gradient|gradient slots|Error|error|timeout" gpurun_out/tests_multirank.log | tail -n 14
run 29541 --steps 20 --warmup 5 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "dp$N exit $?"; cut -c1-300 gpurun_out/bench_dp$N.json; tail -n 2 gpurun_out/bench_dp$N.err
# the 8-GPU per-rank batch (4 pairs per rank) on N ranks: gradient slots on / off, overlapped exchange
run 29544 --global-batch $((4*N)) --steps 30 --warmup 5 > gpurun_out/bench_dp${N}_b4.json 2> gpurun_out/bench_dp${N}_b4.err; echo "dp$N b4/rank exit $?"; cut -c1-200 gpurun_out/bench_dp${N}_b4.json
ADNI_PEER_GRADS=0 run 29545 --global-batch $((4*N)) --steps 30 --warmup 5 > gpurun_out/bench_dp${N}_b4_nccl.json 2> gpurun_out/bench_dp${N}_b4_nccl.err; echo "dp$N b4/rank NCCL gradients exit $?"; cut -c1-200 gpurun_out/bench_dp${N}_b4_nccl.json
for c in 148 592; do ADNI_PEER_GRAD_CTAS=$c run 29547 --global-batch $((4*N)) --steps 30 --warmup 5 > gpurun_out/bench_dp${N}_b4_ctas$c.json 2> gpurun_out/bench_dp${N}_b4_ctas$c.err; echo "dp$N b4/rank $c CTAs exit $?"; cut -c1-200 gpurun_out/bench_dp${N}_b4_ctas$c.json; done
ADNI_OVERLAP_GRADS=1 run 29546 --global-batch $((4*N)) --steps 30 --warmup 5 > gpurun_out/bench_dp${N}_b4_overlap.json 2> gpurun_out/bench_dp${N}_b4_overlap.err; echo "dp$N b4/rank overlap exit $?"; cut -c1-200 gpurun_out/bench_dp${N}_b4_overlap.json
ADNI_OVERLAP_GRADS=1 run 29542 --steps 20 --warmup 5 > gpurun_out/bench_dp${N}_overlap.json 2> gpurun_out/bench_dp${N}_overlap.err; echo "dp$N overlap exit $?"; cut -c1-300 gpurun_out/bench_dp${N}_overlap.json; tail -n 2 gpurun_out/bench_dp${N}_overlap.err
if [ "$2" = "extra" ]; then
  run 29543 --workload mri_r50_160 --steps 5 --warmup 3 > gpurun_out/bench_r50_dp$N.json 2> gpurun_out/bench_r50_dp$N.err; echo "r50 dp$N exit $?"; cut -c1-300 gpurun_out/bench_r50_dp$N.json; tail -n 2 gpurun_out/bench_r50_dp$N.err
fi
