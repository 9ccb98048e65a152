#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/halo.log
for v in "10 0" "10 1" "16 0" "16 1"; do
  set -- $v
  ADNI_HALO_PITCH=$1 ADNI_HALO_BASE=$2 timeout 120 python tools/halo_probe.py >> gpurun_out/halo.log 2>&1
  echo "variant $v exit $?" >> gpurun_out/halo.log
done
cat gpurun_out/halo.log | grep -v Warning | tail -60
