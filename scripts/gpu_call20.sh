#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "wide8 or trace_rays_bit_exact or device_built or million or sah_order" ) > gpurun_out/c20_pytest_first.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest_first.log
tail -30 gpurun_out/c20_pytest_first.log
L=gpurun_out/c20_probe.log; : > $L
echo "== c4: compressed 8-wide (flags 0) vs 4-wide (flags 262144), device-built tree and host SAH tree" >> $L
timeout 400 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --flags 0,262144 --regen 0,14,26 --reps 3 --stats >> $L 2>&1
timeout 400 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --flags 8,262152 --reps 1 --stats >> $L 2>&1
grep -v "^NVIDIA" $L
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.log
tail -8 gpurun_out/c20_pytest.log
