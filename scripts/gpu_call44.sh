#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python scripts/fuzz_parity.py 3000 20000 2>&1 | tail -4 ) > gpurun_out/c44_fuzz.log 2>&1; cat gpurun_out/c44_fuzz.log
