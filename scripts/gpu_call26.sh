#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "wide8 or step_functions or million or device_built or sah_order or binary_and_wide" ) > gpurun_out/c26_pytest_first.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c26_pytest_first.log
tail -4 gpurun_out/c26_pytest_first.log
L=gpurun_out/c26_probe.log; : > $L
for v in pop1 pop2 default pop8; do
  echo "== $v (stack entries fetched per pop round; default = 4; S-block chain shortened in all)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --flags 0,196608 --reps 3 >> $L 2>&1
done
unset TRT_B200_LIB
timeout 300 python scripts/perf_probe.py --workloads c2,c3 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
grep -v "^NVIDIA\|trt_scene_commit" $L
