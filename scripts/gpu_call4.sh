#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/c4_probe.log; : > $L
echo "== wavefront (wide nodes) on the tree workloads" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --variant 1 --reps 3 >> $L 2>&1
TRT_WF_SLOTS=4194304 timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --variant 1 --reps 3 >> $L 2>&1
TRT_WF_SLOTS=262144 timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --variant 1 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c2,c3,c5 --spp 32 --kinds 3 --variant 1 --reps 3 >> $L 2>&1
echo "== megakernel regen sweep c4" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --regen 8,12,16,20,24,28 --reps 2 >> $L 2>&1
echo "== counted walk c4, c2" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 2 --kinds 3,1 --flags 8 --reps 1 --stats >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c2 --spp 8 --kinds 3 --flags 8 --reps 1 --stats >> $L 2>&1
grep -v "^NVIDIA" $L
