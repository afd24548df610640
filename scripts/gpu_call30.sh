#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "first_bounce or progressive or sample_partition or render_multi_on_one" ) > gpurun_out/c30_pytest_first.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c30_pytest_first.log
tail -4 gpurun_out/c30_pytest_first.log
L=gpurun_out/c30_probe.log; : > $L
for v in default noprefetch; do
  echo "== $v (first-bounce records: one 32-byte record per path, 256-bit load; default also prefetches the pixel's next record into L2)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --flags 0,16 --reps 3 >> $L 2>&1
  timeout 300 python scripts/ncu_one.py c2 500 >> $L 2>&1
done
grep -v "^NVIDIA\|trt_scene_commit" $L
