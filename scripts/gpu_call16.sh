#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c16_pytest.log
L=gpurun_out/c16_probe.log; : > $L
echo "== first-bounce pass (default)" >> $L
timeout 300 python scripts/perf_probe.py --workloads c3,c2,c5 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c3 --spp 256 --kinds 3 --reps 2 >> $L 2>&1
echo "== inline primary rays (flag 16)" >> $L
timeout 300 python scripts/perf_probe.py --workloads c3,c2,c5 --spp 64 --kinds 3 --flags 16 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --flags 16 --reps 3 >> $L 2>&1
tail -6 gpurun_out/c16_pytest.log; grep -v "^NVIDIA" $L
