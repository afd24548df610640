"""Small renders through every kernel family (flat list, reference tree, SAH on wide and binary nodes, device-built tree,
step-wise schedule, wavefront, AOV / trace / sample-batch hooks): the command compute-sanitizer is pointed at."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tiny_raytracer_b200 as trt

ctx = trt.Context(0)
S = trt.scenes
cases = [(S.cornell_spheres(48, 32, 3, 8), [trt.BVH_FLAT, trt.BVH_REFERENCE, trt.BVH_SAH]),
         (S.glass_cornell(40, 24, 2, 12), [trt.BVH_AUTO]),
         (S.random_mixed(300, 5, 48, 32, 2, 8), [trt.BVH_SAH, trt.BVH_REFERENCE, trt.BVH_LBVH_DEVICE]),
         (S.random_spheres(6000, 48, 27, 2, 6), [trt.BVH_SAH, trt.BVH_LBVH_DEVICE]),
         (S.final_scene(60, 34, 3, 20), [trt.BVH_AUTO]),                                   # first-bounce pass + sign-selected staged wide nodes
         (S.random_mixed(3000, 17, 48, 27, 2, 10), [trt.BVH_SAH, trt.BVH_LBVH_DEVICE])]    # wide nodes in global memory: 256-bit loads, batched pops; compressed 8-wide nodes (TRT_WIDE8=1 below)
os.environ["TRT_WIDE8"] = "1"
for desc, kinds in cases:
    cam = trt.make_camera(**desc.camera)
    w, h = desc.camera["width"], desc.camera["height"]
    for kind in kinds:
        sc = trt.Scene.from_desc(desc, kind)
        for flags in (0, 4, 12 << 8, 1, 16, 32, (12 << 8) | (1 << 16), (12 << 8) | (3 << 16), (12 << 8) | 0x40000, 8 | (12 << 8)):   # + inline / forced first-bounce pass, step functions, 4-wide walk beside wide-8 nodes, walked counters
            for variant in (trt.VARIANT_MEGAKERNEL, trt.VARIANT_WAVEFRONT):
                if variant == trt.VARIANT_WAVEFRONT and flags:
                    continue
                p = trt.make_params(w, h, desc.spp, desc.max_bounces, desc.background, desc.seed, variant=variant, flags=flags)
                lin, u8, st = sc.render(ctx, cam, p)
                assert np.isfinite(lin).all()
        p = trt.make_params(w, h, desc.spp, desc.max_bounces, desc.background, desc.seed)
        sc.render_aov(ctx, cam, p, True, True)
        rng = np.random.default_rng(1)
        o = rng.uniform(-50, 50, (5000, 3)).astype(np.float32); d = rng.normal(size=(5000, 3)).astype(np.float32)
        sc.trace_rays(ctx, o, d, 0.001)
        sc.close()
        print(desc.name, kind, "ok", flush=True)
print("sanitize probe done")
