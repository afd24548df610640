#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c41_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c41_pytest.log
tail -4 gpurun_out/c41_pytest.log
( timeout 600 python scripts/fuzz_parity.py 300 11000 2>&1 | tail -3 ) > gpurun_out/c41_fuzz.log 2>&1; cat gpurun_out/c41_fuzz.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c41_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c41_smoke.log; tail -2 gpurun_out/c41_smoke.log
timeout 300 python scripts/ncu_one.py c5 64 > gpurun_out/c41_probe.log 2>&1
timeout 300 python scripts/ncu_one.py c3 128 >> gpurun_out/c41_probe.log 2>&1
timeout 300 python scripts/perf_probe.py --workloads c1 --spp 16 --kinds 3 --reps 5 >> gpurun_out/c41_probe.log 2>&1
grep -v "^NVIDIA\|trt_scene_commit" gpurun_out/c41_probe.log
timeout 900 python bench.py --workload c5 --spp 512 --steps 2 --warmup 3 --cpu-spp 8 > gpurun_out/r2l_bench_c5_512spp.json 2> gpurun_out/r2l_bench_c5_512spp.err
grep -o '"value": [0-9.]*' gpurun_out/r2l_bench_c5_512spp.json | head -2; grep -o '"linear_max_ulp[^,]*' gpurun_out/r2l_bench_c5_512spp.json
