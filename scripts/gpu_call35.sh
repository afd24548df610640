#!/usr/bin/env bash
# final state of round 2 on an 8-GPU box: C3 under torchrun (NCCL reduce, parity at N = 8) and through the C ABI (trt_multi_render)
N=8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c3 --steps 10 --warmup 3 > gpurun_out/r2k_bench_c3_n8.json 2> gpurun_out/r2k_bench_c3_n8.err; echo "rc=$?" >> gpurun_out/r2k_bench_c3_n8.err
timeout 600 python bench.py --gpus $N --native-multi --workload c3 --steps 10 --warmup 3 > gpurun_out/r2k_bench_c3_native_n8.json 2> gpurun_out/r2k_bench_c3_native_n8.err; echo "rc=$?" >> gpurun_out/r2k_bench_c3_native_n8.err
for f in gpurun_out/r2k_*_n8.json; do echo "$f: $(grep -o '"value": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"ms_per_step": [0-9.]*' $f | head -1) $(grep -o '"linear_max_ulp[^,]*' $f | head -1)"; done
grep -h "rc=" gpurun_out/r2k_*_n8.err | sort | uniq -c
