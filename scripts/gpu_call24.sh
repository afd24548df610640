#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/c24_probe.log; : > $L
for v in default prim_na na_mb5 hot2 hot4; do
  echo "== $v" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --reps 3 >> $L 2>&1
done
unset TRT_B200_LIB
echo "== default: regen sweep, 4-wide 128-bit loads for reference (flags 65536)" >> $L
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --regen 12,16,24,28 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --flags 65536 --reps 3 >> $L 2>&1
grep -v "^NVIDIA\|trt_scene_commit" $L
