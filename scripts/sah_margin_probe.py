"""How many samples of a random-sphere scene differ between the SAH walk and the oracle, all bounces.
Usage: python scripts/sah_margin_probe.py [n_spheres] [spp] [depth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tiny_raytracer_b200 as trt
from oracle import pyoracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 2
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 12
desc = trt.scenes.random_spheres(n, 960, 540, spp, depth)
ctx = trt.Context(0)
cam = trt.make_camera(**desc.camera); ocam = orc.make_camera(**desc.camera)
osc = orc.Scene.from_desc(desc)
p = trt.make_params(960, 540, spp, depth, desc.background, desc.seed)
op = orc.make_params(960, 540, spp, depth, desc.background, desc.seed)
t0 = time.time(); oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True); t_or = time.time() - t0
for kind in (trt.BVH_SAH, trt.BVH_LBVH_DEVICE):
    sc = trt.Scene.from_desc(desc, kind)
    prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
    same = (nr == onr) & (rad.view(np.uint32) == orad.view(np.uint32)).all(1) & (prim == oprim) & (t.view(np.uint32) == ot.view(np.uint32))
    print("lib=%s n=%d kind=%d samples=%d path mismatches=%d primary-hit mismatches=%d ray-count mismatches=%d (oracle %.1fs)" % (
        os.environ.get("TRT_B200_LIB", "default"), n, kind, same.size, int((~same).sum()), int((prim != oprim).sum()), int((nr != onr).sum()), t_or), flush=True)
