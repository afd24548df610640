#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/c47_probe.log; : > $L
for v in default sstack8 sstack12; do
  echo "== $v (step-wise loop, trees in global memory: the first 8 / 12 stack entries of every thread in shared memory, default carve-out; default = local memory only, max-L1 carve-out)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --reps 3 >> $L 2>&1
done
grep -v "^NVIDIA\|trt_scene_commit" $L
