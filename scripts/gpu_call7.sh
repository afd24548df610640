#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.log
L=gpurun_out/c7_probe.log; : > $L
echo "== c4: one path per lane (flag 16), regen sweep" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --flags 16 --regen 0,16,24 --reps 2 >> $L 2>&1
echo "== c4: two paths per lane, regen sweep (swap 4)" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --regen 8,12,16,20,24,28,31 --reps 2 >> $L 2>&1
echo "== c4: two paths per lane, swap sweep at regen 20 / 24" >> $L
for sw in 1 2 8 12; do timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --regen 20,24 --flags $((sw<<16)) --reps 2 >> $L 2>&1; done
echo "== c4 sah tree" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 1 --regen 0,24 --reps 2 >> $L 2>&1
echo "== c2 / c3-tree: T-then-S (default) vs step-wise two paths vs one path" >> $L
timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --regen 0,12,16,20,24,28 --reps 2 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --regen 16,24 --flags 16 --reps 2 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c3 --spp 32 --kinds 1 --regen 0,16,24 --reps 2 >> $L 2>&1
tail -4 gpurun_out/c7_pytest.log; grep -v "^NVIDIA" $L
