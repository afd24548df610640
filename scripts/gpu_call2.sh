#!/usr/bin/env bash
# round-2 call 2: A/B of flat-kernel variants (split phase B, 112-byte records) and traversal-stack variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
for v in "" nosplit7 nosplit6 split6; do
  echo "== flat variant: ${v:-main}" >> gpurun_out/c2_probe.log
  TRT_B200_LIB=${v:+$PWD/build/variants/libtrt_$v.so} timeout 300 python scripts/perf_probe.py --workloads c3,c5 --spp 64 --kinds 3 --reps 4 >> gpurun_out/c2_probe.log 2>&1
done
for v in "" stack0 smem12 smem16 smem24; do
  echo "== stack variant: ${v:-main}" >> gpurun_out/c2_probe.log
  TRT_B200_LIB=${v:+$PWD/build/variants/libtrt_$v.so} timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --reps 3 >> gpurun_out/c2_probe.log 2>&1
  TRT_B200_LIB=${v:+$PWD/build/variants/libtrt_$v.so} timeout 300 python scripts/perf_probe.py --workloads c2 --spp 64 --kinds 3 --reps 3 >> gpurun_out/c2_probe.log 2>&1
done
tail -4 gpurun_out/c2_pytest.log; grep -v "^NVIDIA" gpurun_out/c2_probe.log
