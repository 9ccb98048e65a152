cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_dp4.json 2> gpurun_out/bench_dp4.err; echo "dp4 exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_dp4.json') if l.startswith('{')][-1])
print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms','e2e',round(d['e2e']['value'],1),d['gradient_exchange'],d['sync_bn_exchange'],d['parity']['first_step_loss'],d['parity']['first_step_logits_checksum'])
PY
tail -3 gpurun_out/bench_dp4.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 > gpurun_out/bench_ref_dp4.json 2> gpurun_out/bench_ref_dp4.err; echo "reference arm under torchrun exit $?"; cut -c1-200 gpurun_out/bench_ref_dp4.json
