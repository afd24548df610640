#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c36_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c36_pytest.log
tail -5 gpurun_out/c36_pytest.log
( timeout 600 python scripts/fuzz_parity.py 200 7000 2>&1 | tail -3 ) > gpurun_out/c36_fuzz.log 2>&1; cat gpurun_out/c36_fuzz.log
L=gpurun_out/c36_probe.log; : > $L
for v in default noitempf; do
  echo "== $v (T-then-S loops: the next work item is reserved one closest-hit query early; noitempf = before)" >> $L
  if [ $v = default ]; then unset TRT_B200_LIB; else export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so; fi
  timeout 300 python scripts/perf_probe.py --workloads c3 --spp 128 --kinds 3 --reps 4 >> $L 2>&1
  timeout 300 python scripts/perf_probe.py --workloads c3 --spp 512 --kinds 3 --reps 2 >> $L 2>&1
  timeout 300 python scripts/perf_probe.py --workloads c1,c2,c5 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
  timeout 300 python scripts/perf_probe.py --workloads c1 --spp 16 --kinds 3 --reps 5 >> $L 2>&1
done
grep -v "^NVIDIA\|trt_scene_commit" $L
