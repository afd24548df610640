#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c18_pytest.log
L=gpurun_out/c18_probe.log; : > $L
S1=65536; S2=131072; S3=196608
echo "== c4 device-built tree: step functions 1/2/3 x regen" >> $L
timeout 400 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3 --flags $S1,$S2,$S3 --regen 0,14,26 --reps 3 >> $L 2>&1
echo "== c4 host SAH tree" >> $L
timeout 400 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 1 --flags $S1,$S2,$S3 --reps 3 >> $L 2>&1
echo "== small scenes: auto / inline / pass" >> $L
timeout 300 python scripts/perf_probe.py --workloads c3,c2,c5 --spp 64 --kinds 3 --flags 0,16,32 --reps 3 >> $L 2>&1
tail -8 gpurun_out/c18_pytest.log; grep -v "^NVIDIA" $L
