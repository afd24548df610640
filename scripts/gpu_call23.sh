#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/c23_probe.log; : > $L
export TRT_COMMIT_TIMING=
for mb in none 32 48 64 96; do
  echo "== L2 access-policy window on the wide nodes: TRT_L2_PERSIST_MB=$mb" >> $L
  if [ $mb = none ]; then unset TRT_L2_PERSIST_MB; else export TRT_L2_PERSIST_MB=$mb; fi
  timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --reps 3 >> $L 2>&1
done
unset TRT_L2_PERSIST_MB
for v in noalloc evlast; do
  echo "== node loads with $v" >> $L
  export TRT_B200_LIB=$PWD/build/variants/libtrt_$v.so
  timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --reps 3 >> $L 2>&1
done
grep -v "^NVIDIA" $L
