#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python scripts/ncu_one.py c2 64 > gpurun_out/c43_plain_c2.log 2>&1 && $NCU -k regex:k_render_mega -s 2 -c 1 -f -o gpurun_out/r2l_tree_c2_64spp python scripts/ncu_one.py c2 64 > gpurun_out/c43_ncu_c2.log 2>&1
$NCU -k regex:k_primary -s 2 -c 1 -f -o gpurun_out/r2l_primary_c2_64spp python scripts/ncu_one.py c2 64 > gpurun_out/c43_ncu_c2p.log 2>&1
cat gpurun_out/c43_plain_c2.log; ls -la gpurun_out/r2l_*.ncu-rep
