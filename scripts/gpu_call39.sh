#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/tail_probe.py c5 short 128,512 > gpurun_out/c39_tail_c5.log 2>&1
timeout 300 python scripts/tail_probe.py c3 short 1024 > gpurun_out/c39_tail_c3.log 2>&1
timeout 300 python scripts/tail_probe.py c2 short 500 > gpurun_out/c39_tail_c2.log 2>&1
grep -hv "^NVIDIA" gpurun_out/c39_tail_c5.log gpurun_out/c39_tail_c3.log gpurun_out/c39_tail_c2.log
