#!/bin/bash
# Validation + A/B of the opt-in CTA-pair conv engine (csrc/conv_igemm_2cta.cu); first results: profiles/r01_2cta_ab.json
#   gpurun --timeout 900 -- 'bash tools/gpu_2cta.sh'
# 1. conv parity tests with the engine forced on for every N = 256 tile (fprop Cout 256/512, dgrad Cin 256/512);
# 2. if green: the bench with and without it (same box, back to back).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ADNI_IGEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/tests_2cta.log 2>&1
rc=$?
echo "2cta conv tests exit $rc"; tail -n 5 gpurun_out/tests_2cta.log
if [ $rc -eq 0 ]; then
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1cta.log 2>&1
  ADNI_IGEMM_2CTA=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2cta.log 2>&1
  for f in bench_1cta bench_2cta; do echo $f; tail -n 1 gpurun_out/$f.log | cut -c1-220; done
fi
