"""Writes tests/golden/workload_counters.json: the oracle's event counts (reference-topology tree) per workload on a
bounded sample, used by bench.py for the algorithmic bytes / ops per ray when no live oracle run is available (N > 1).
Run here (CPU): python scripts/make_workload_counters.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tiny_raytracer_b200 as trt
from oracle import pyoracle as orc

SAMPLE = {"c1": 16, "c1s": 16, "c2": 4, "c3": 4, "c4": 1, "c5": 1}
out = {}
for wl, spp in SAMPLE.items():
    desc = trt.scenes.WORKLOADS[wl]()
    c = desc.camera
    t0 = time.time()
    sc = orc.Scene.from_desc(desc); cam = orc.make_camera(**c)
    p = orc.make_params(c["width"], c["height"], spp, desc.max_bounces, desc.background, desc.seed, spp_total=spp)
    _, _, st, _ = sc.render(cam, p)
    out[wl] = {"rays": int(st.rays), "paths": int(st.paths), "node_visits": int(st.node_visits), "sphere_tests": int(st.sphere_tests),
               "quad_tests": int(st.quad_tests), "hits": int(st.hits), "sample_spp": spp, "tree": sc.tree_stats(),
               "workload": "%s %dx%d depth %d" % (desc.name, c["width"], c["height"], desc.max_bounces)}
    r = out[wl]["rays"]
    print(wl, "%.1fs" % (time.time() - t0), "rays/path %.2f V %.1f Ps %.2f Pq %.2f H %.2f" % (r / st.paths, st.node_visits / r, st.sphere_tests / r, st.quad_tests / r, st.hits / r), flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "workload_counters.json"), "w"), indent=1)
