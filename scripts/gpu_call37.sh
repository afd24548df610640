#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c37_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c37_pytest.log
tail -5 gpurun_out/c37_pytest.log
( timeout 600 python scripts/fuzz_parity.py 300 9000 2>&1 | tail -3 ) > gpurun_out/c37_fuzz.log 2>&1; cat gpurun_out/c37_fuzz.log
L=gpurun_out/c37_probe.log; : > $L
for v in default flatb0; do
  echo "== $v (flat list phase B: pre-rejects in an inner loop until a lane has a survivor, exact halves together; flatb0 = before)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c3,c5,c1 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
  timeout 300 python scripts/perf_probe.py --workloads c3 --spp 256 --kinds 3 --reps 2 >> $L 2>&1
done
grep -v "^NVIDIA\|trt_scene_commit" $L
