#!/usr/bin/env bash
# scripts/gpu_scale.sh N : the N-GPU evidence of round 2 on an N-GPU box — C3 under torchrun (NCCL reduce) with parity at N,
# C3 through the C ABI (trt_multi_render: threads + NVLink peer reduce-scatter + tone map), the C++ example, and C5 both ways.
N=$1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus_n$N.txt
run_torch() { # workload out extra...
  local wl=$1 out=$2; shift 2
  if [ "$N" = 1 ]; then timeout 1500 python bench.py --gpus 1 --workload $wl "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err
  else timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $wl "$@" > gpurun_out/$out.json 2> gpurun_out/$out.err; fi
  echo "rc=$?" >> gpurun_out/$out.err
}
if [ "${SKIP_C3:-0}" != 1 ]; then
  run_torch c3 r2e_bench_c3_n$N --steps 5 --warmup 3
  timeout 900 python bench.py --gpus $N --native-multi --workload c3 --steps 5 --warmup 3 > gpurun_out/r2e_bench_c3_native_n$N.json 2> gpurun_out/r2e_bench_c3_native_n$N.err; echo "rc=$?" >> gpurun_out/r2e_bench_c3_native_n$N.err
  g++ -std=c++17 -O2 -Iinclude examples/cornell_multi.cpp -o /tmp/cornell_multi -Ltiny-raytracer_b200 -ltrt_b200 -Wl,-rpath,$PWD/tiny-raytracer_b200
  timeout 600 /tmp/cornell_multi $N 1920 1080 1024 > gpurun_out/r2e_cornell_multi_cpp_n$N.json 2> gpurun_out/r2e_cornell_multi_cpp_n$N.err; echo "rc=$?" >> gpurun_out/r2e_cornell_multi_cpp_n$N.err
fi
run_torch c5 r2e_bench_c5_n$N --steps 2 --warmup 3 --cpu-spp 8
if [ "$N" != 1 ]; then
  timeout 900 python bench.py --gpus $N --native-multi --workload c5 --steps 2 --warmup 3 > gpurun_out/r2e_bench_c5_native_n$N.json 2> gpurun_out/r2e_bench_c5_native_n$N.err; echo "rc=$?" >> gpurun_out/r2e_bench_c5_native_n$N.err
fi
for f in gpurun_out/r2e_*_n$N.json; do echo "$f: $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"linear_max_ulp[^,]*' $f | head -1) $(grep -o '"efficiency_device": [0-9.]*' $f)"; done
grep -h "rc=" gpurun_out/r2e_*_n$N.err | sort | uniq -c
