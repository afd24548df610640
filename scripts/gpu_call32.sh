#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c32_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c32_pytest.log
tail -5 gpurun_out/c32_pytest.log
( timeout 600 python scripts/fuzz_parity.py 300 2>&1 | tail -3 ) > gpurun_out/c32_fuzz.log 2>&1; cat gpurun_out/c32_fuzz.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c32_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c32_smoke.log; tail -2 gpurun_out/c32_smoke.log
timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2k_bench_c2.json').read().strip().splitlines()[-1])
print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('bound'), d['roofline'].get('frac'), d['parity'].get('linear_max_ulp'), d.get('gpu_launches'))
P
