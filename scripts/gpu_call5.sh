#!/usr/bin/env bash
# round-2 call 5 (2 GPUs): N-GPU parity on hardware, native multi-GPU through the C ABI, torchrun bench with parity at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c5_gpus.txt; nvidia-smi topo -m >> gpurun_out/c5_gpus.txt 2>&1
( timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or render_multi or cpp_multi" ) > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
g++ -std=c++17 -O2 -Iinclude examples/cornell_multi.cpp -o /tmp/cornell_multi -Ltiny-raytracer_b200 -ltrt_b200 -Wl,-rpath,$PWD/tiny-raytracer_b200
timeout 300 /tmp/cornell_multi 2 1920 1080 256 > gpurun_out/c5_cornell_multi_n2.json 2> gpurun_out/c5_cornell_multi_n2.err; echo "rc=$?" >> gpurun_out/c5_cornell_multi_n2.err
TRT_MULTI_STAGED=1 timeout 300 /tmp/cornell_multi 2 1920 1080 64 > gpurun_out/c5_cornell_multi_n2_staged.json 2>> gpurun_out/c5_cornell_multi_n2.err
timeout 600 python bench.py --gpus 2 --native-multi --steps 3 --warmup 3 > gpurun_out/c5_bench_native_n2.json 2> gpurun_out/c5_bench_native_n2.err; echo "rc=$?" >> gpurun_out/c5_bench_native_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err; echo "rc=$?" >> gpurun_out/c5_bench_n2.err
tail -5 gpurun_out/c5_pytest.log; cat gpurun_out/c5_cornell_multi_n2.json gpurun_out/c5_cornell_multi_n2_staged.json; tail -3 gpurun_out/c5_cornell_multi_n2.err
head -c 1500 gpurun_out/c5_bench_native_n2.json; echo; tail -3 gpurun_out/c5_bench_native_n2.err; head -c 600 gpurun_out/c5_bench_n2.json; echo; tail -3 gpurun_out/c5_bench_n2.err
