#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/halo_debug_probe.py > gpurun_out/halo_dbg.log 2>&1
echo "exit $?"; grep -v Warn gpurun_out/halo_dbg.log | tail -30
