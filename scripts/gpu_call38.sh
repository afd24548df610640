#!/usr/bin/env bash
# final state of round 2 (HEAD): GPU suite, smoke, the default bench command with its ncu launch list, and the bench lines of the two configurations not re-measured since r2c (C1, C5 share)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c38_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c38_pytest.log
tail -4 gpurun_out/c38_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c38_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/c38_smoke.log; tail -2 gpurun_out/c38_smoke.log
timeout 600 python bench.py > gpurun_out/r2l_bench_default.json 2> gpurun_out/r2l_bench_default.err; echo "rc=$?" >> gpurun_out/r2l_bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; echo "rc=$?" >> gpurun_out/r2l_bench_ref.err
timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/r2l_bench_c1.json 2> gpurun_out/r2l_bench_c1.err
timeout 900 python bench.py --workload c5 --spp 512 --steps 2 --warmup 3 --cpu-spp 8 > gpurun_out/r2l_bench_c5_512spp.json 2> gpurun_out/r2l_bench_c5_512spp.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c38_ncu_bench.log 2>&1
for f in default ref c1 c5_512spp; do python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r2l_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('parity') or {}).get('linear_max_ulp'), d.get('gpu_launches'), d.get('impl'))
except Exception as e: print('$f', 'ERR', e)
P
done
grep -c "k_render_mega_flat" gpurun_out/r2l_launches_bench_c3.csv
