#!/usr/bin/env bash
# round-2 call 10: ncu evidence for the final kernels at the BENCH configurations (launch list of the bench command, --set full
# captures for traffic.json), and the megakernel / wavefront pair on C3 and C5 (north star item 2) with the current kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_ncu_bench.log 2>&1
python scripts/ncu_one.py c3 1024 > gpurun_out/r2d_plain_c3.log 2>&1 && $NCU -k regex:k_render_mega -s 2 -c 1 -f -o gpurun_out/r2d_flat_c3_1024spp python scripts/ncu_one.py c3 1024 > gpurun_out/r2d_ncu_c3.log 2>&1
python scripts/ncu_one.py c2 500 > gpurun_out/r2d_plain_c2.log 2>&1 && $NCU -k regex:k_render_mega -s 2 -c 1 -f -o gpurun_out/r2d_tree_c2_500spp python scripts/ncu_one.py c2 500 > gpurun_out/r2d_ncu_c2.log 2>&1
python scripts/ncu_one.py c4 64 > gpurun_out/r2d_plain_c4.log 2>&1 && $NCU -k regex:k_render_mega -s 2 -c 1 -f -o gpurun_out/r2d_tree_c4_64spp python scripts/ncu_one.py c4 64 > gpurun_out/r2d_ncu_c4.log 2>&1
python scripts/ncu_one.py c5 64 > gpurun_out/r2d_plain_c5.log 2>&1 && $NCU -k regex:k_render_mega -s 2 -c 1 -f -o gpurun_out/r2d_flat_c5_64spp python scripts/ncu_one.py c5 64 > gpurun_out/r2d_ncu_c5.log 2>&1
# wavefront variant, same workloads at 16 spp: a few (shade, extend) pairs from the middle of the render
python scripts/ncu_one.py c3 16 3 0 1 > gpurun_out/r2d_plain_wf3.log 2>&1 && $NCU -k regex:k_wf_ -s 40 -c 4 -f -o gpurun_out/r2d_wf_c3_16spp python scripts/ncu_one.py c3 16 3 0 1 > gpurun_out/r2d_ncu_wf3.log 2>&1
python scripts/ncu_one.py c5 4 3 0 1 > gpurun_out/r2d_plain_wf5.log 2>&1 && $NCU -k regex:k_wf_ -s 40 -c 4 -f -o gpurun_out/r2d_wf_c5_4spp python scripts/ncu_one.py c5 4 3 0 1 > gpurun_out/r2d_ncu_wf5.log 2>&1
cat gpurun_out/r2d_plain_*.log | grep -v "^{" ; ls -la gpurun_out/r2d_*.ncu-rep gpurun_out/r2d_launches_bench_c3.csv
