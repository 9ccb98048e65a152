#!/bin/bash
# session-3 check 3: even-kernel 'same' padding (new tests only), then the stock PyTorch + cuDNN comparator
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q --timeout 200 -p no:cacheprovider -k "even_kernel or fmf or early_fusion or fusion_ops" > gpurun_out/tests_r1e.log 2>&1
echo "tests exit $?"; tail -n 6 gpurun_out/tests_r1e.log | cut -c1-300
timeout 110 python tests/cudnn_comparator.py 32 > gpurun_out/cudnn_comparator.log 2>&1
echo "comparator exit $?"; tail -n 2 gpurun_out/cudnn_comparator.log | cut -c1-900
