#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c29_pytest.log
tail -5 gpurun_out/c29_pytest.log
timeout 300 python scripts/e2e_probe.py c4 8 > gpurun_out/c29_e2e_c4.log 2>&1
timeout 300 python scripts/e2e_probe.py c4 8 1 >> gpurun_out/c29_e2e_c4.log 2>&1
grep -v "^NVIDIA" gpurun_out/c29_e2e_c4.log
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err
for f in c4 c3; do python - <<P
import json
d=json.loads(open('gpurun_out/r2j_bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('bound'), d['roofline'].get('frac'), d['parity'].get('linear_max_ulp'), d.get('gpu_launches'))
P
done
NCU="ncu --set full --clock-control none --import-source on -k regex:k_render_mega -s 2 -c 1"
python scripts/ncu_one.py c4 64 > gpurun_out/c29_plain_c4.log 2>&1 && $NCU -f -o gpurun_out/r2j_tree_c4_64spp python scripts/ncu_one.py c4 64 > gpurun_out/c29_ncu_c4.log 2>&1
python scripts/ncu_one.py c2 500 > gpurun_out/c29_plain_c2.log 2>&1 && $NCU -f -o gpurun_out/r2j_tree_c2_500spp python scripts/ncu_one.py c2 500 > gpurun_out/c29_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_render_mega|k_primary|k_finalize" --csv --log-file gpurun_out/r2j_launches_c2_500spp.csv python scripts/ncu_one.py c2 500 > gpurun_out/c29_ncu_c2_list.log 2>&1
cat gpurun_out/c29_plain_c*.log; ls -la gpurun_out/r2j*
