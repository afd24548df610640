#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c15_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c15_smoke.log
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/c15_bench_c4.json 2> gpurun_out/c15_bench_c4.err
tail -3 gpurun_out/c15_pytest.log; cat gpurun_out/c15_smoke.log; grep -o '"value": [0-9.]*' gpurun_out/c15_bench_c4.json | head -3
