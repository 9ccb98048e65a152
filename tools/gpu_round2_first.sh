#!/bin/bash
# First GPU call of the next round (one B200):  gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh'
#  1. the whole GPU suite + smoke at the committed state
#  2. the default bench line (-> gpurun_out/bench_default.json)
#  3. the opt-in variants written at the end of round 1, each validated by the existing bit-exactness tests before it
#     is timed: packed-compare stem pooling (ADNI_POOL_STREAM=2), CTA-pair conv engine (ADNI_IGEMM_2CTA=1)
#  4. training from .nii.gz files through StagedLoader (cold / cached epochs)
# 2-GPU items (overlapped gradient all-reduce): gpurun --gpus 2 -- 'bash tools/gpu_overlap_dp.sh 2'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_all.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/tests_all.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_default.json
bash tools/gpu_pool_packed.sh
bash tools/gpu_2cta.sh
timeout 500 python tools/bench_staged_e2e.py > gpurun_out/staged_e2e.json 2> gpurun_out/staged_e2e.err; echo "staged e2e exit $?"; cat gpurun_out/staged_e2e.json
