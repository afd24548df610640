#!/usr/bin/env bash
# ncu --set full of the 10^6-sphere kernel: 4-wide nodes with 256-bit loads (flags NO_WIDE8) and compressed 8-wide nodes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:k_render_mega -s 2 -c 1"
python scripts/ncu_one.py c4 2 3 262144 > gpurun_out/c21_plain_w4.log 2>&1 && $NCU -f -o gpurun_out/r2g_tree_c4_2spp_wide4_ld256 python scripts/ncu_one.py c4 2 3 262144 > gpurun_out/c21_ncu_w4.log 2>&1
python scripts/ncu_one.py c4 2 3 0 > gpurun_out/c21_plain_w8.log 2>&1 && $NCU -f -o gpurun_out/r2g_tree_c4_2spp_wide8 python scripts/ncu_one.py c4 2 3 0 > gpurun_out/c21_ncu_w8.log 2>&1
cat gpurun_out/c21_plain_*.log; tail -2 gpurun_out/c21_ncu_*.log; ls -la gpurun_out/r2g*.ncu-rep
