#!/usr/bin/env bash
# scripts/sass_fn.sh <lib.so> <mangled-substring>  -> the SASS of the first matching function
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{f=(index($0,pat)>0)} f'
