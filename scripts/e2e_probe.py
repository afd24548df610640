"""Where does the end-to-end time of a big scene go? commit (host) / upload (H2D + device builds) / render + image back, wall clock."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiny_raytracer_b200 as trt
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = int(sys.argv[3]) if len(sys.argv) > 3 else trt.BVH_AUTO
desc = trt.scenes.WORKLOADS[wl]()
w, h = desc.camera["width"], desc.camera["height"]
ctx = trt.Context(0)
cam = trt.make_camera(**desc.camera)
sc = trt.Scene.from_desc(desc, kind)
lin = np.empty(w * h * 3, np.float32); u8 = np.empty(w * h * 3, np.uint8)
p = trt.make_params(w, h, spp, desc.max_bounces, desc.background, desc.seed, spp_total=spp)
for rep in range(4):
    t0 = time.perf_counter(); sc.commit(kind); t1 = time.perf_counter()
    sc.upload(ctx); torch.cuda.synchronize(); t2 = time.perf_counter()
    _, _, st = sc.render(ctx, cam, p, out_linear=lin, out_rgb8=u8); t3 = time.perf_counter()
    print("%s kind=%d spp=%d: commit %.1f ms, upload %.1f ms (device build %.1f ms), render call %.1f ms (kernel %.1f ms, tone map %.2f ms)" % (
        wl, kind, spp, 1e3 * (t1 - t0), 1e3 * (t2 - t1), sc.build_times()["device_build_ms"], 1e3 * (t3 - t2), st.device_ms, st.finalize_ms), flush=True)
