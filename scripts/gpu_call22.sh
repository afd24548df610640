#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "wide8 or first_bounce or step_functions" ) > gpurun_out/c22_pytest_first.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c22_pytest_first.log
tail -5 gpurun_out/c22_pytest_first.log
L=gpurun_out/c22_probe.log; : > $L
for v in default mb3 mb5 mb6; do
  echo "== $v (CTAs per SM of the tree megakernels; c4 and c2)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --reps 3 >> $L 2>&1
  timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --flags 0,16 --reps 3 >> $L 2>&1
done
grep -v "^NVIDIA" $L
