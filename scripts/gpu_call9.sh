#!/usr/bin/env bash
# round-2 call 9: bench lines of all five BASELINE configs on one GPU, with the final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err; echo "rc=$?" >> gpurun_out/r2c_bench_c3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err
timeout 600 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/r2c_bench_c1.json 2> gpurun_out/r2c_bench_c1.err
timeout 900 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err
timeout 1500 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err; echo "rc=$?" >> gpurun_out/r2c_bench_c4.err
timeout 1500 python bench.py --workload c4 --bvh sah --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/r2c_bench_c4_sah.json 2> gpurun_out/r2c_bench_c4_sah.err
timeout 1500 python bench.py --workload c5 --spp 512 --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/r2c_bench_c5_512spp.json 2> gpurun_out/r2c_bench_c5.err; echo "rc=$?" >> gpurun_out/r2c_bench_c5.err
for f in gpurun_out/r2c_bench_*.json; do echo $f; head -c 400 $f; echo; done; tail -2 gpurun_out/r2c_bench_*.err
