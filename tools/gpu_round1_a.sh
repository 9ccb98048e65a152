#!/bin/bash
# First GPU bring-up: kernel-level parity tests, each group in its own process (a trapped kernel poisons the context).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
i=0
for g in "tests/test_gpu_kernels.py" "tests/test_gpu_conv.py -k direct_conv" "tests/test_gpu_conv.py -k tc_fprop" "tests/test_gpu_conv.py -k tc_dgrad" "tests/test_gpu_conv.py -k tc_wgrad"; do
  i=$((i+1))
  timeout 420 python -m pytest $g -m gpu -q --timeout 100 -p no:cacheprovider > gpurun_out/a_$i.log 2>&1
  echo "group $i ($g) exit $?" >> gpurun_out/a_summary.txt
  tail -n 3 gpurun_out/a_$i.log >> gpurun_out/a_summary.txt
done
cat gpurun_out/a_summary.txt
