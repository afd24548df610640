"""Quick device-timed probe of the render kernels on the BASELINE workloads at reduced spp (development aid)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tiny_raytracer_b200 as trt

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="c3,c2,c4,c5")
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--kinds", default="0,1")
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--c4n", type=int, default=1_000_000)
ap.add_argument("--flags", default="0", help="comma list of extra trt_params.flags values to sweep (2 = no tail split, 16 = inline primary rays, 32 = first-bounce pass, n << 16 = step function n)")
ap.add_argument("--stats", action="store_true", help="also print the event counters of the best run")
ap.add_argument("--regen", default="0", help="comma list of regen_min overrides (0 = auto)")
a = ap.parse_args()

ctx = trt.Context(0)
print(ctx.name, ctx.sm_count, flush=True)
for wl in a.workloads.split(","):
    desc = trt.scenes.WORKLOADS[wl]() if wl != "c4" else trt.scenes.random_spheres(a.c4n)
    if wl == "c5":
        desc = desc.with_image(1920, 1080)
    w, h = desc.camera["width"], desc.camera["height"]
    cam = trt.make_camera(**desc.camera)
    for kind in [int(k) for k in a.kinds.split(",")]:
        t0 = time.time()
        sc = trt.Scene.from_desc(desc, kind)
        t_build = time.time() - t0
        sc.upload(ctx)
        acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        for regen, xflags in [(int(x), int(f)) for f in a.flags.split(",") for x in a.regen.split(",")]:
            p = trt.make_params(w, h, a.spp, desc.max_bounces, desc.background, desc.seed, spp_total=a.spp, variant=a.variant, flags=((regen & 255) << 8) | xflags)
            best = None
            for r in range(a.reps):
                acc.zero_(); torch.cuda.synchronize()
                st = sc.render_accumulate(ctx, cam, p, acc.data_ptr())
                if best is None or st.device_ms < best.device_ms:
                    best = st
            print("%-3s kind=%d variant=%d spp=%d flags=%d regen=%d  build %.2fs  nodes %d  %.2f ms  %.1f Mpaths/s  %.1f Mrays/s  rays/path %.2f" % (
                wl, kind, a.variant, a.spp, xflags, regen, t_build, sc.bvh_info()["nodes"], best.device_ms, best.paths / best.device_ms / 1e3,
                best.rays / best.device_ms / 1e3, best.rays / max(best.paths, 1)), flush=True)
            if a.stats:
                r = max(best.rays, 1)
                print("    events per ray: nodes %.2f  sphere tests %.2f  quad tests %.2f" % (best.node_visits / r, best.sphere_tests / r, best.quad_tests / r), flush=True)
        sc.close()
