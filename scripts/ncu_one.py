"""One render of a BASELINE workload at reduced spp for profiling: two warm-up launches, then the launch to capture
(run under: ncu --set full --clock-control none --import-source on -k regex:k_render_mega -s 2 -c 1 -o gpurun_out/X python scripts/ncu_one.py c3 32).  argv: workload spp [bvh kind] [flags] [variant]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tiny_raytracer_b200 as trt
wl, spp = sys.argv[1], int(sys.argv[2])
kind = int(sys.argv[3]) if len(sys.argv) > 3 else trt.BVH_AUTO
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
desc = trt.scenes.WORKLOADS[wl]()
w, h = desc.camera["width"], desc.camera["height"]
ctx = trt.Context(0)
cam = trt.make_camera(**desc.camera)
sc = trt.Scene.from_desc(desc, kind); sc.upload(ctx)
acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:0"); torch.cuda.synchronize()
p = trt.make_params(w, h, spp, desc.max_bounces, desc.background, desc.seed, spp_total=spp, flags=flags, variant=variant)
for i in range(3):
    acc.zero_(); torch.cuda.synchronize()
    st = sc.render_accumulate(ctx, cam, p, acc.data_ptr())
print("%s spp=%d kind=%d: %.2f ms, %d rays, %.1f Mrays/s" % (wl, spp, kind, st.device_ms, st.rays, st.rays / st.device_ms / 1e3))
