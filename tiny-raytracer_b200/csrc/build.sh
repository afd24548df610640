#!/usr/bin/env bash
# Builds libtrt_b200.so in-tree for sm_100a. IEEE-faithful arithmetic on both sides: no FMA contraction, correctly
# rounded division and square root, denormals kept.
set -euo pipefail
cd "$(dirname "$0")"
CUDA=${CUDA_HOME:-/usr/local/cuda}
NVCC=${NVCC:-$CUDA/bin/nvcc}
CXX=${CXX:-g++}
OUT=${TRT_OUT:-../libtrt_b200.so}
DEVFLAGS="${TRT_EXTRA_DEVFLAGS:-} -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false"
HOSTFLAGS="-std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -Wall -pthread"
# the five translation units are independent: compile them side by side
pids=()
$NVCC $DEVFLAGS -Xcompiler -fPIC,-ffp-contract=off ${TRT_PTXAS_V:+-Xptxas -v} -c trt_kernels.cu -o trt_kernels.o & pids+=($!)
$NVCC $DEVFLAGS -Xcompiler -fPIC,-ffp-contract=off ${TRT_PTXAS_V:+-Xptxas -v} -c trt_wavefront.cu -o trt_wavefront.o & pids+=($!)
$NVCC $DEVFLAGS -Xcompiler -fPIC -c trt_lbvh.cu -o trt_lbvh.o & pids+=($!)
$CXX $HOSTFLAGS -I$CUDA/include -c trt_capi.cpp -o trt_capi.o & pids+=($!)
$CXX $HOSTFLAGS -c trt_scene.cpp -o trt_scene.o & pids+=($!)
for pid in "${pids[@]}"; do wait "$pid"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT trt_kernels.o trt_wavefront.o trt_lbvh.o trt_capi.o trt_scene.o -Xcompiler -pthread -cudart shared
echo "built $(realpath $OUT)"
