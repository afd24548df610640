"""Tail-split tuning of the persistent megakernel: device time of the Cornell 1080p frame at 128 spp (one GPU's share of 8) and
256 spp for a few settings of TRT_TAIL_PX / TRT_TAIL_DIV / TRT_TAIL_MIN (read by launch_render_mega at every launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tiny_raytracer_b200 as trt
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc = trt.scenes.WORKLOADS[wl]()
w, h = desc.camera["width"], desc.camera["height"]
ctx = trt.Context(0)
cam = trt.make_camera(**desc.camera)
sc = trt.Scene.from_desc(desc, trt.BVH_AUTO); sc.upload(ctx)
acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:0"); torch.cuda.synchronize()
settings = [(2, 32, 4), (1, 32, 4), (0.5, 32, 4), (4, 32, 4), (2, 16, 4), (2, 64, 4), (2, 64, 2), (2, 128, 1), (1, 64, 2), (1, 16, 8), (3, 32, 4), (1.5, 32, 4)]
if len(sys.argv) > 2 and sys.argv[2] == "short": settings = [(2, 32, 4), (3, 32, 4), (2, 32, 4), (3, 32, 4), (4, 32, 4), (1, 32, 4)]
for spp in ((128, 256) if not (len(sys.argv) > 3) else tuple(int(x) for x in sys.argv[3].split(","))):
    p = trt.make_params(w, h, spp, desc.max_bounces, desc.background, desc.seed, spp_total=spp)
    for px, div, mn in settings:
        os.environ["TRT_TAIL_PX"] = str(px); os.environ["TRT_TAIL_DIV"] = str(div); os.environ["TRT_TAIL_MIN"] = str(mn)
        best = 1e9
        for r in range(4):
            acc.zero_(); torch.cuda.synchronize()
            st = sc.render_accumulate(ctx, cam, p, acc.data_ptr())
            best = min(best, st.device_ms)
        print("%s spp=%d tail_px=%s div=%d min=%d: %.3f ms" % (wl, spp, px, div, mn, best), flush=True)
