#!/usr/bin/env bash
# final state of round 2 on a 2-GPU box: the two-device tests, C3 under torchrun (NCCL reduce) and through the C ABI (trt_multi_render), C2 and C4 under torchrun
N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus or multi" ) > gpurun_out/c34_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c34_pytest_multi.log
tail -4 gpurun_out/c34_pytest_multi.log
run_torch() { local wl=$1 out=$2; shift 2
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "rc=$?" >> gpurun_out/$out.err; }
run_torch c3 r2k_bench_c3_n2 --steps 5 --warmup 3
timeout 900 python bench.py --gpus $N --native-multi --workload c3 --steps 5 --warmup 3 > gpurun_out/r2k_bench_c3_native_n2.json 2> gpurun_out/r2k_bench_c3_native_n2.err; echo "rc=$?" >> gpurun_out/r2k_bench_c3_native_n2.err
run_torch c2 r2k_bench_c2_n2 --steps 5 --warmup 3 --no-cpu-baseline
run_torch c4 r2k_bench_c4_n2 --steps 3 --warmup 3 --cpu-spp 2
for f in gpurun_out/r2k_*_n2.json; do echo "$f: $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"ms_per_step": [0-9.]*' $f | head -1) $(grep -o '"linear_max_ulp[^,]*' $f | head -1)"; done
grep -h "rc=" gpurun_out/r2k_*_n2.err | sort | uniq -c
