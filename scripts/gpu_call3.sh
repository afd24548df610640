#!/usr/bin/env bash
# round-2 call 3: ncu --set full captures (with source) of the three kernels: flat (C3), tree in shared memory (C2), tree in L2 (C4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:k_render_mega -s 2 -c 1"
python scripts/ncu_one.py c3 32 > gpurun_out/c3_plain.log 2>&1 && $NCU -f -o gpurun_out/r2a_flat_c3_32spp python scripts/ncu_one.py c3 32 > gpurun_out/c3_ncu.log 2>&1
python scripts/ncu_one.py c2 16 > gpurun_out/c3_plain2.log 2>&1 && $NCU -f -o gpurun_out/r2a_tree_c2_16spp python scripts/ncu_one.py c2 16 > gpurun_out/c3_ncu2.log 2>&1
python scripts/ncu_one.py c4 2 > gpurun_out/c3_plain4.log 2>&1 && $NCU -f -o gpurun_out/r2a_tree_c4_2spp python scripts/ncu_one.py c4 2 > gpurun_out/c3_ncu4.log 2>&1
cat gpurun_out/c3_plain*.log; tail -3 gpurun_out/c3_ncu*.log; ls -la gpurun_out/*.ncu-rep
