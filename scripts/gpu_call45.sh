#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/c45_probe.log; : > $L
echo "== final scene: step-wise loop thresholds with the first-bounce pass forced (flags 32), and without (flags 16)" >> $L
timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --flags 32 --regen 0,12,16,20,24,28 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --flags 16 --regen 0,20,28 --reps 3 >> $L 2>&1
grep -v "^NVIDIA\|trt_scene_commit" $L
