#!/usr/bin/env bash
# round-2 call 1: full GPU suite on the new code, a short bench line, baseline probes of the tree workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpus.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/c1_bench_c3.json 2> gpurun_out/c1_bench_c3.err
echo "bench rc=$?" >> gpurun_out/c1_bench_c3.err
timeout 600 python scripts/perf_probe.py --workloads c3,c2,c5 --spp 64 --kinds 3 > gpurun_out/c1_probe.log 2>&1
timeout 600 python scripts/perf_probe.py --workloads c4 --spp 8 --kinds 3,1 >> gpurun_out/c1_probe.log 2>&1
tail -5 gpurun_out/c1_pytest.log; cat gpurun_out/c1_bench_c3.json | head -c 3000; cat gpurun_out/c1_probe.log
