"""Differential fuzzing on the GPU box: random mixed worlds (spheres, quads, all four materials, coplanar and duplicate
primitives) of random sizes, random cameras and depths, every BVH kind, whole paths against the oracle bit for bit.
Usage: python scripts/fuzz_parity.py [n_scenes] [first_seed]   -> prints one line per failure, a summary at the end."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tiny_raytracer_b200 as trt
from oracle import pyoracle as orc

n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = trt.Context(0)
rng = np.random.default_rng(seed0)
fails = checked = 0
t_start = time.time()
for k in range(n_scenes):
    n = int(rng.choice([1, 2, 3, 5, 17, 40, 64, 65, 130, 700, 3000]))
    w, h = int(rng.integers(9, 90)), int(rng.integers(5, 60))
    spp, depth = int(rng.integers(1, 5)), int(rng.choice([1, 2, 7, 30]))
    desc = trt.scenes.random_mixed(n, seed0 + k, w, h, spp, depth)
    cam_args = dict(desc.camera)
    cam_args["defocus_angle"] = float(rng.choice([0.0, 0.5, 3.0]))
    cam_args["position"] = tuple(float(x) for x in rng.uniform(-25, 25, 3))
    cam_args["vertical_fov"] = float(rng.uniform(20, 100))
    desc = trt.scenes.SceneDesc(desc.name, desc.materials, desc.prim_type, desc.prim_data, desc.prim_mat, cam_args, spp, depth,
                                tuple(float(x) for x in rng.random(3)), int(rng.integers(1, 1 << 40)))
    ocam = orc.make_camera(**cam_args); cam = trt.make_camera(**cam_args)
    osc = orc.Scene.from_desc(desc)
    op = orc.make_params(w, h, spp, depth, desc.background, desc.seed)
    oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True)
    olin, ou8, ost, _ = osc.render(ocam, op)
    n = desc.n_prims
    if n == 0:
        continue
    # arbitrary rays too, a third of them degenerate (zero / tiny / axis-aligned components, origins on integer planes, far away)
    nr_ = 3000
    ro = rng.uniform(-14, 14, (nr_, 3)).astype(np.float32); rd = rng.normal(size=(nr_, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True).astype(np.float32)
    rd[:150, int(rng.integers(3))] = 0.0; rd[150:300, int(rng.integers(3))] = np.float32(-0.0); rd[300:400, int(rng.integers(3))] = np.float32(1e-30)
    rd[400:500] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 100)] * np.float32(rng.choice([-1.0, 1.0]))
    ro[500:800] = np.round(ro[500:800]); ro[800:900] *= np.float32(1e6); ro[900:1000, 2] = np.float32(rng.integers(-3, 4))
    oprim_r, ot_r = osc.trace_rays(ro, rd, 0.001, use_bvh=True)
    kinds = [trt.BVH_REFERENCE, trt.BVH_SAH, trt.BVH_AUTO] + ([trt.BVH_LBVH_DEVICE] if n >= 2 else []) + ([trt.BVH_FLAT] if n <= 64 else [])
    for kind in kinds:
        sc = trt.Scene.from_desc(desc, kind)
        prim_r, t_r = sc.trace_rays(ctx, ro, rd, 0.001)
        checked += 1
        if not (np.array_equal(prim_r, oprim_r) and np.array_equal(t_r.view(np.uint32), ot_r.view(np.uint32))):
            fails += 1
            print("FAIL scene %d (n=%d seed=%d) kind=%d trace_rays: prim %d t %d" % (k, n, seed0 + k, kind, int((prim_r != oprim_r).sum()),
                  int((t_r.view(np.uint32) != ot_r.view(np.uint32)).sum())), flush=True)
        for flags in (0, 4, 9 << 8, (3 << 8) | 4, 16, (9 << 8) | 16):      # T-then-S, binary nodes, step-wise loop on wide / binary nodes; inline primary rays
            p = trt.make_params(w, h, spp, depth, desc.background, desc.seed, flags=flags | 2)
            prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
            lin, u8, st = sc.render(ctx, cam, p)
            ok = (np.array_equal(prim, oprim) and np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and np.array_equal(nr, onr)
                  and np.array_equal(rad.view(np.uint32), orad.view(np.uint32)) and st.rays == ost.rays
                  and np.array_equal(lin.view(np.uint32), olin.view(np.uint32)))
            checked += 1
            if not ok:
                fails += 1
                print("FAIL scene %d (n=%d %dx%d spp=%d depth=%d seed=%d) kind=%d flags=%d: prim %d t %d nr %d rad %d rays %d/%d lin %d" % (
                    k, n, w, h, spp, depth, seed0 + k, kind, flags, int((prim != oprim).sum()), int((t.view(np.uint32) != ot.view(np.uint32)).sum()),
                    int((nr != onr).sum()), int((rad.view(np.uint32) != orad.view(np.uint32)).any(1).sum()), st.rays, ost.rays,
                    int((lin.view(np.uint32) != olin.view(np.uint32)).sum())), flush=True)
        if kind in (trt.BVH_REFERENCE, trt.BVH_AUTO) and k % 4 == 0:
            # the wavefront variant traces the same paths through queues in HBM; only the f64 summation order differs (<= 1 ulp)
            p = trt.make_params(w, h, spp, depth, desc.background, desc.seed, variant=trt.VARIANT_WAVEFRONT)
            lin, u8, st = sc.render(ctx, cam, p)
            ulp = np.abs(lin.view(np.int32).astype(np.int64) - olin.view(np.int32).astype(np.int64)).max() if lin.size else 0
            checked += 1
            if not (st.rays == ost.rays and st.paths == ost.paths and ulp <= 1):
                fails += 1
                print("FAIL scene %d (n=%d seed=%d) kind=%d wavefront: rays %d/%d ulp %d" % (k, n, seed0 + k, kind, st.rays, ost.rays, int(ulp)), flush=True)
        sc.close()
print("fuzz: %d scenes, %d (scene, kind, flags) renders checked, %d failures, %.0f s" % (n_scenes, checked, fails, time.time() - t_start))
sys.exit(1 if fails else 0)
