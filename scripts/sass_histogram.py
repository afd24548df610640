"""SASS opcode histogram per kernel of libtrt_b200.so (static counts from `cuobjdump -sass`): the evidence for the claims
about the instruction mix — packed fp32 (FFMA2 / FADD2 / FMUL2), 3-input FMNMX3, uniform-datapath constant loads (LDCU), local
memory (LDL / STL = spills or an in-memory stack), MUFU, and that no tensor-core / TMA instruction is expected on this path.
Usage: python scripts/sass_histogram.py [lib.so] [out.json]   (default: the in-tree library -> profiles/r2_sass_histogram.json)"""
import collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tiny-raytracer_b200", "libtrt_b200.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2_sass_histogram.json")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FMNMX3", "FMNMX", "FSETP", "MUFU", "LDCU", "LDC", "LDS", "STS", "LDG", "STG", "LDL", "STL",
         "ATOMG", "RED", "DADD", "F2F", "IMAD", "LOP3", "SHFL", "VOTE", "BRA", "BSSY", "BSYNC", "WARPSYNC", "CALL", "UTMALDG", "UTCMMA", "HMMA", "TCGEN05"]
res, cur, hist = {}, None, None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        hist = res.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and hist is not None:
        hist[m.group(1)] += 1
        hist["_total"] += 1
summary = {}
for k, h in sorted(res.items()):
    summary[k] = {"total": h["_total"], **{w: h[w] for w in WATCH if h[w]},
                  "other_top": {o: c for o, c in h.most_common(40) if o not in WATCH and o != "_total"}}
json.dump({"library": os.path.relpath(lib, ROOT), "how": "cuobjdump -sass, static instruction counts per kernel (not execution counts)", "kernels": summary},
          open(out, "w"), indent=1)
for k, v in summary.items():
    print("%-60s total %6d  FFMA2 %4d FADD2 %4d FMUL2 %4d FMNMX3 %4d LDCU %4d LDL %4d STL %4d MUFU %4d" % (
        k[:60], v["total"], v.get("FFMA2", 0), v.get("FADD2", 0), v.get("FMUL2", 0), v.get("FMNMX3", 0), v.get("LDCU", 0), v.get("LDL", 0), v.get("STL", 0), v.get("MUFU", 0)))
