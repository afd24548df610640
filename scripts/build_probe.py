"""Host vs device BVH build time and the traversal speed of the resulting trees (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tiny_raytracer_b200 as trt
ctx = trt.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
desc = trt.scenes.random_spheres(n, 1920, 1080, 8, 50)
w, h = 1920, 1080
cam = trt.make_camera(**desc.camera)
for kind, name in ((trt.BVH_SAH, "host SAH"), (trt.BVH_REFERENCE, "host reference"), (trt.BVH_LBVH_DEVICE, "device LBVH")):
    t0 = time.time(); sc = trt.Scene.from_desc(desc, kind); t_commit = time.time() - t0
    t0 = time.time(); sc.upload(ctx); ctx.synchronize(); t_up = time.time() - t0
    acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:0"); torch.cuda.synchronize()
    p = trt.make_params(w, h, 4, desc.max_bounces, desc.background, desc.seed, spp_total=4)
    best = None
    for r in range(2):
        acc.zero_(); torch.cuda.synchronize()
        st = sc.render_accumulate(ctx, cam, p, acc.data_ptr())
        best = st if best is None or st.device_ms < best.device_ms else best
    print("%-15s n=%d commit %.3fs (host tree %.1f ms) upload %.3fs (device build %.2f ms) depth %d  render 4spp %.1f ms %.1f Mrays/s" % (
        name, n, t_commit, sc.build_times()["host_commit_ms"], t_up, sc.build_times()["device_build_ms"], sc.bvh_info()["max_depth"],
        best.device_ms, best.rays / best.device_ms / 1e3), flush=True)
    sc.close()
