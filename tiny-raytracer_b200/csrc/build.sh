#!/usr/bin/env bash
# Builds libtrt_b200.so in-tree for sm_100a. IEEE-faithful arithmetic on both sides: no FMA contraction, correctly
# rounded division and square root, denormals kept. The translation units are listed in sources.txt.
#   TRT_OUT=<path>            output library (default ../libtrt_b200.so)
#   TRT_OBJDIR=<dir>          object directory (default: this directory) — lets A/B variants build side by side
#   TRT_EXTRA_DEVFLAGS="..."  extra nvcc flags (e.g. -DTRT_DEBUG_BOUNDS), TRT_PTXAS_V=1 adds -Xptxas -v
#   TRT_EXTRA_DEFS="-D..."    macros for BOTH the device and the host translation units (layout constants)
set -euo pipefail
cd "$(dirname "$0")"
CUDA=${CUDA_HOME:-/usr/local/cuda}
NVCC=${NVCC:-$CUDA/bin/nvcc}
CXX=${CXX:-g++}
OUT=${TRT_OUT:-../libtrt_b200.so}
OBJ=${TRT_OBJDIR:-.}
mkdir -p "$OBJ"
DEVFLAGS="${TRT_EXTRA_DEVFLAGS:-} ${TRT_EXTRA_DEFS:-} -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false"
HOSTFLAGS="${TRT_EXTRA_DEFS:-} -std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -Wall -pthread"
# the translation units are independent: compile them side by side
pids=(); objs=()
while read -r src kind; do
    case "$src" in ''|'#'*) continue;; esac
    o="$OBJ/${src%.*}.o"; objs+=("$o")
    if [ "$kind" = cu ]; then
        $NVCC $DEVFLAGS -Xcompiler -fPIC,-ffp-contract=off,-pthread ${TRT_PTXAS_V:+-Xptxas -v} -c "$src" -o "$o" & pids+=($!)
    else
        $CXX $HOSTFLAGS -I$CUDA/include -c "$src" -o "$o" & pids+=($!)
    fi
done < sources.txt
for pid in "${pids[@]}"; do wait "$pid"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT "${objs[@]}" -Xcompiler -pthread -cudart shared
echo "built $(realpath $OUT)"
