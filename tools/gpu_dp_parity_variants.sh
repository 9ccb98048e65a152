cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 tools/dp_parity.py "${@:2}"; }
ADNI_FLAT_1X1=0 run 29571 --workload pet_mri_fusion_r18 --volume 64 --depth 18 --per-rank 2 > gpurun_out/dp_flat0.log 2>&1; echo "flat off, per-rank 2: exit $?"; grep -E "gradients:|PARITY|grad rel" gpurun_out/dp_flat0.log | head -8
run 29572 --workload pet_mri_fusion_r18 --volume 64 --depth 18 --per-rank 2 > gpurun_out/dp_flat1.log 2>&1; echo "flat on, per-rank 2: exit $?"; grep -E "gradients:|PARITY|grad rel" gpurun_out/dp_flat1.log | head -8
run 29573 --workload pet_mri_fusion_r18 --volume 64 --depth 18 --per-rank 4 > gpurun_out/dp_b8.log 2>&1; echo "flat on, per-rank 4: exit $?"; grep -E "gradients:|PARITY|grad rel|loss sharded|logits rel" gpurun_out/dp_b8.log | head -10
run 29574 --workload mri_r50_160 --volume 48 --depth 50 --per-rank 3 > gpurun_out/dp_r50_b6.log 2>&1; echo "r50 per-rank 3: exit $?"; grep -E "gradients:|PARITY|grad rel|loss sharded|logits rel" gpurun_out/dp_r50_b6.log | head -10
