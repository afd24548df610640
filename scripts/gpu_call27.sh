#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/drain_probe.py c3 > gpurun_out/c27_drain_c3.log 2>&1
grep -v "^NVIDIA" gpurun_out/c27_drain_c3.log
timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/r2i_bench_c4.json 2> gpurun_out/r2i_bench_c4.err
timeout 900 python bench.py --workload c4 --bvh sah --steps 3 --warmup 3 --cpu-spp 2 > gpurun_out/r2i_bench_c4_sah.json 2> gpurun_out/r2i_bench_c4_sah.err
timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err
for f in c4 c4_sah c2; do python - <<P
import json
d=json.loads(open('gpurun_out/r2i_bench_$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('bound'), d['roofline'].get('frac'), d.get('parity'), d.get('gpu_launches'), d['config'].get('kernel'))
P
done
