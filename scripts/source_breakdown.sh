#!/usr/bin/env bash
# scripts/source_breakdown.sh <report.ncu-rep> <mangled kernel substring> <rays of the launch> <out.txt> [lib.so]
# Executed warp instructions of one kernel attributed to source lines (scripts/sass_line_profile.py), innermost line and
# one level of inlining context.
set -euo pipefail
rep=$1; pat=$2; rays=$3; out=$4; lib=${5:-tiny-raytracer_b200/libtrt_b200.so}
tmp=$(mktemp -d)
ncu -i "$rep" --page source --csv --print-source sass > "$tmp/sass.csv"
(cd "$tmp" && cuobjdump -xelf all "$OLDPWD/$lib" > /dev/null)
: > "$tmp/lines.txt"
for c in "$tmp"/*.cubin; do nvdisasm -gi "$c" 2>/dev/null | awk -v pat="$pat" '/^\s*\.text\./{f=(index($0,pat)>0)} /^\s*\.section/{ if (index($0,".text.")==0) f=0 } f' >> "$tmp/lines.txt"; done
{ python scripts/sass_line_profile.py "$tmp/sass.csv" "$tmp/lines.txt" "$rays"; echo "---- with one level of inlining context ----"; DEPTH=2 python scripts/sass_line_profile.py "$tmp/sass.csv" "$tmp/lines.txt" "$rays"; } > "$out"
rm -rf "$tmp"
head -40 "$out"
