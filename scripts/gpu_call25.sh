#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.log
tail -6 gpurun_out/c25_pytest.log
NCU="ncu --set full --clock-control none --import-source on -k regex:k_render_mega -s 2 -c 1"
python scripts/ncu_one.py c4 2 > gpurun_out/c25_plain_c4.log 2>&1 && $NCU -f -o gpurun_out/r2h_tree_c4_2spp python scripts/ncu_one.py c4 2 > gpurun_out/c25_ncu_c4.log 2>&1
cat gpurun_out/c25_plain_c4.log
L=gpurun_out/c25_probe.log; : > $L
timeout 300 python scripts/perf_probe.py --workloads c3,c2,c5 --spp 64 --kinds 3 --reps 3 >> $L 2>&1
timeout 300 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 --reps 3 >> $L 2>&1
grep -v "^NVIDIA\|trt_scene_commit" $L
