#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.log
L=gpurun_out/c8_probe.log; : > $L
timeout 300 python scripts/perf_probe.py --workloads c3,c2,c5 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c3 --spp 128 --kinds 3 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 1,0 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c2,c4 --spp 8 --kinds 3 --variant 1 --reps 2 >> $L 2>&1
tail -4 gpurun_out/c8_pytest.log; grep -v "^NVIDIA" $L
