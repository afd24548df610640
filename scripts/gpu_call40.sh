#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/c40_probe.log; : > $L
for v in default geomearly default geomearly; do
  echo "== $v (flat list: hit geometry computed with the material for every hit lane (geomearly, round-1 form) or later for the scattering lanes only (default))" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/ncu_one.py c5 64 >> $L 2>&1
  timeout 300 python scripts/ncu_one.py c3 128 >> $L 2>&1
done
grep -v "^NVIDIA\|trt_scene_commit" $L
