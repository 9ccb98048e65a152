#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_parity.py > gpurun_out/dp_parity.log 2>&1
echo "dp parity exit $?"; tail -n 6 gpurun_out/dp_parity.log
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/models.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/models.log
