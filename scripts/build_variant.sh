#!/usr/bin/env bash
# A/B builds of the same ABI: scripts/build_variant.sh <name> "<device flags>" ["<defs for host and device>"]
#   -> build/variants/libtrt_<name>.so (select with TRT_B200_LIB=...). build/ is git-ignored but travels to the GPU box.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name=$1; flags=${2:-}; defs=${3:-}
mkdir -p "$ROOT/build/variants/obj_$name"
TRT_OUT="$ROOT/build/variants/libtrt_$name.so" TRT_OBJDIR="$ROOT/build/variants/obj_$name" TRT_EXTRA_DEVFLAGS="$flags" TRT_EXTRA_DEFS="$defs" \
    bash "$ROOT/tiny-raytracer_b200/csrc/build.sh" 2>&1 | tail -1
