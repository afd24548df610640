#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/e2e_probe.py c4 8 > gpurun_out/c28_e2e_c4.log 2>&1
timeout 300 python scripts/e2e_probe.py c4 8 1 >> gpurun_out/c28_e2e_c4.log 2>&1
grep -v "^NVIDIA" gpurun_out/c28_e2e_c4.log
timeout 600 python scripts/tail_probe.py c3 > gpurun_out/c28_tail_c3.log 2>&1
grep -v "^NVIDIA" gpurun_out/c28_tail_c3.log
