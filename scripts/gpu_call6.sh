#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
L=gpurun_out/c6_probe.log; : > $L
for v in "" mb3 mb5 mb6; do
  echo "== variant: ${v:-main (sphere shortcuts, 4 CTAs/SM)}" >> $L
  TRT_B200_LIB=${v:+$PWD/build/variants/libtrt_$v.so} timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --reps 3 >> $L 2>&1
  TRT_B200_LIB=${v:+$PWD/build/variants/libtrt_$v.so} timeout 300 python scripts/perf_probe.py --workloads c2,c5 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
done
tail -4 gpurun_out/c6_pytest.log; grep -v "^NVIDIA" $L
