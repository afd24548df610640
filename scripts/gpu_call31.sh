#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c31_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c31_pytest.log
tail -5 gpurun_out/c31_pytest.log
L=gpurun_out/c31_probe.log; : > $L
for v in default sign; do
  echo "== $v (wide nodes in shared memory: near / far plane pairs through sign-dependent addresses; default = min/max form, sign = sign-dependent addresses)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3,0 --flags 0,16 --reps 3 >> $L 2>&1
  timeout 300 python scripts/ncu_one.py c2 500 >> $L 2>&1
done
grep -v "^NVIDIA\|trt_scene_commit" $L
