#!/bin/bash
# Final 8-GPU confirmation:  gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_dp8_final.sh'
#   exchange kernels against NCCL + the sharded step against the CPU oracle at 8 ranks, then the default bench line.
N=8
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 tools/dp_parity.py --workload pet_mri_fusion_r18 --volume 64 --depth 18 --per-rank 1 > gpurun_out/dp_parity8.log 2>&1; echo "dp_parity at 8 ranks exit $?"; grep -E "peer all-reduce|peer gradient|gradients:|PARITY|loss sharded|logits rel|bit-identical|timeout|Error" gpurun_out/dp_parity8.log | head -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_dp8_final.json 2> gpurun_out/bench_dp8_final.err; echo "dp8 exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_dp8_final.json') if l.startswith('{')][-1])
print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms','e2e',round(d['e2e']['value'],1),d['gradient_exchange'],'|',d['sync_bn_exchange'],'| first loss',d['parity']['first_step_loss'],d['parity']['first_step_logits_checksum'])
PY
