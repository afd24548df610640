"""Generates tests/golden/golden_small.npz from the CPU oracle (oracle/oracle.hpp, MATH_DET). The reference itself cannot
be run here or anywhere on Linux (Rust + macOS Metal) and its RNG is unseeded, so these are ORACLE outputs,
committed to pin the oracle against regressions and to give the GPU tests fixtures that need no oracle run.
Run here (CPU): python tests/golden/make_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import tiny_raytracer_b200 as trt
from oracle import pyoracle as orc

CASES = {
    "cornell": lambda S: S.cornell(48, 36, 2, 20),
    "cornell_spheres": lambda S: S.cornell_spheres(48, 36, 2, 20),
    "five_spheres": lambda S: S.five_spheres(40, 30, 2, 10),
    "five_quads": lambda S: S.five_quads(40, 30, 2, 10),
    "final_scene": lambda S: S.final_scene(40, 22, 2, 50),
    "glass_cornell": lambda S: S.glass_cornell(32, 18, 2, 100),
    "random_mixed": lambda S: S.random_mixed(120, 31, 32, 24, 2, 12),
}

if __name__ == "__main__":
    out = {}
    for name, mk in CASES.items():
        d = mk(trt.scenes); c = d.camera
        sc = orc.Scene.from_desc(d); cam = orc.make_camera(**c)
        p = orc.make_params(c["width"], c["height"], d.spp, d.max_bounces, d.background, d.seed)
        prim, t, rad, nr = sc.render_aov(cam, p, True, True)
        lin, u8, st, _ = sc.render(cam, p)
        out[name + "/prim"] = prim; out[name + "/t"] = t; out[name + "/rad"] = rad; out[name + "/nr"] = nr.astype(np.uint16)
        out[name + "/lin"] = lin; out[name + "/u8"] = u8
        out[name + "/counts"] = np.array([st.paths, st.rays, st.node_visits, st.sphere_tests, st.quad_tests, st.hits], np.uint64)
        out[name + "/ranks"] = sc.leaf_ranks()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_small.npz"), **out)
    print("wrote", sum(v.nbytes for v in out.values()), "bytes (uncompressed)")
