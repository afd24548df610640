"""Where does the fixed cost of a megakernel launch go? Device time of the Cornell 1080p frame (BASELINE config 3) as a function of
spp, with and without the tail split, for a straight-line fit ms = a * spp + c (c is what limits 8-GPU scaling)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tiny_raytracer_b200 as trt
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc = trt.scenes.WORKLOADS[wl]()
w, h = desc.camera["width"], desc.camera["height"]
ctx = trt.Context(0)
cam = trt.make_camera(**desc.camera)
sc = trt.Scene.from_desc(desc, trt.BVH_AUTO); sc.upload(ctx)
acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:0"); torch.cuda.synchronize()
for flags in (0, 2):
    xs, ys = [], []
    for spp in (0, 1, 2, 4, 8, 16, 32, 64, 128, 256):
        p = trt.make_params(w, h, spp, desc.max_bounces, desc.background, desc.seed, spp_total=max(spp, 1), flags=flags)
        best = 1e9
        for r in range(4):
            acc.zero_(); torch.cuda.synchronize()
            st = sc.render_accumulate(ctx, cam, p, acc.data_ptr())
            best = min(best, st.device_ms)
        print("%s flags=%d spp=%4d  %.3f ms  %.4f ms/spp" % (wl, flags, spp, best, best / max(spp, 1)), flush=True)
        if spp >= 32: xs.append(spp); ys.append(best)
    a, c = np.polyfit(xs, ys, 1)
    print("%s flags=%d fit over spp >= 32: ms = %.4f * spp + %.3f" % (wl, flags, a, c), flush=True)
