"""Attribute executed warp instructions of one kernel to source lines.
Inputs: (1) the SASS page of an ncu report:   ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
        (2) the kernel's disassembly with line info:   nvdisasm -g <cubin>  (cut to the kernel's .text section)
Usage: python scripts/sass_line_profile.py sass.csv lines.txt [rays]   -> table sorted by executed warp instructions."""
import csv, re, sys
from collections import defaultdict

sass_csv, lines_txt = sys.argv[1], sys.argv[2]
rays = float(sys.argv[3]) if len(sys.argv) > 3 else None

# with `nvdisasm -gi` every instruction is preceded by its inline chain, innermost frame first; DEPTH = how many frames,
# counted from the kernel's own frame, form the key (0 = innermost line only)
import os
DEPTH = int(os.environ.get("DEPTH", "0"))
addr_line = {}
chain, fresh = [], True
for ln in open(lines_txt):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append("%s:%s" % (m.group(1).split("/")[-1].replace("trt_", "").replace(".cuh", "").replace(".cu", ""), m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and chain:
        fresh = True
        if DEPTH == 0:
            key = chain[0]
        else:
            outer = chain[::-1][1:1 + DEPTH]          # skip the kernel's own frame
            key = " < ".join(outer[::-1]) if outer else chain[-1]
        addr_line[int(m.group(1), 16)] = ((key, 0), m.group(2).strip())

rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
per_line = defaultdict(lambda: [0, 0, 0])
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= it or not r[ia]:
        continue
    a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    if base is None:
        base = a
    off = a - base
    ex, th = int(r[ie] or 0), int(r[it] or 0)
    key = addr_line.get(off, (("?", 0), ""))[0]
    per_line[key][0] += ex; per_line[key][1] += th; per_line[key][2] += 1
    total += ex
print("total warp instructions executed: %d" % total)
div = rays / 32.0 if rays else None
print("%-40s %6s %14s %7s %8s %s" % ("file:line", "sass", "warp inst", "share", "thr/inst", "per 32 rays" if div else ""))
for key, (ex, th, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0]):
    if ex == 0:
        continue
    print("%-40s %6d %14d %6.2f%% %8.1f %s" % (key[0], n, ex, 100.0 * ex / total, th / ex, ("%.1f" % (ex / div)) if div else ""))
