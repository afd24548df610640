#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TRT_B200_LIB=$PWD/build/variants/libtrt_bounds.so timeout 600 python scripts/sanitize_probe.py > gpurun_out/c42_bounds.log 2>&1; echo "rc=$?" >> gpurun_out/c42_bounds.log
tail -12 gpurun_out/c42_bounds.log
