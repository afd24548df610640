#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json: Mrays/s & Mpaths/s, Cornell 1080p).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one whole render of the workload: every rank renders its share of the samples-per-pixel of every pixel
(spp partitioning, strong scaling: total work fixed), the f64 accumulation buffers are reduced onto rank 0 with NCCL and
rank 0 runs the tone-map epilogue. `value` is whole-job Mrays/s (ray segments = iterations of the bounce loop,
sampler/cpu.rs:47) device-timed with CUDA events, max over ranks. One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

ALGO = {"node_bytes": 32, "sphere_bytes": 16, "quad_bytes": 64, "hit_bytes": 32, "path_bytes": 16,
        "node_ops": 30, "sphere_ops": 27, "quad_ops": 49, "hit_ops": 75}   # SURVEY.md §8(d) per-event figures


def workload_desc(name, spp=None):
    import tiny_raytracer_b200 as trt
    desc = trt.scenes.WORKLOADS[name]()
    return desc.with_render(spp=spp) if spp else desc


def workload_string(desc):
    c = desc.camera
    return "%s %dx%d, %d spp, max depth %d, %d primitives" % (desc.name, c["width"], c["height"], desc.spp, desc.max_bounces, desc.n_prims)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def golden_counters(workload):
    p = os.path.join(ROOT, "tests", "golden", "workload_counters.json")
    if os.path.exists(p):
        return json.load(open(p)).get(workload)
    return None


def per_ray_from_counts(c):
    rays = max(int(c["rays"]), 1)
    v, ps, pq, h, paths = c["node_visits"] / rays, c["sphere_tests"] / rays, c["quad_tests"] / rays, c["hits"] / rays, c["paths"] / rays
    return {"V": v, "Ps": ps, "Pq": pq, "H": h,
            "bytes": ALGO["node_bytes"] * v + ALGO["sphere_bytes"] * ps + ALGO["quad_bytes"] * pq + ALGO["hit_bytes"] * h + ALGO["path_bytes"] * paths,
            "ops": ALGO["node_ops"] * v + ALGO["sphere_ops"] * ps + ALGO["quad_ops"] * pq + ALGO["hit_ops"] * h}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for i, nm in enumerate(names):
                if f[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(power)) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_render_sample(desc, spp, n_threads=0, sample_offset=0):
    """The reference's CPU implementation of the path = the oracle port (the Rust crate cannot be built: SURVEY.md F8)."""
    from oracle import pyoracle as orc
    c = desc.camera
    cam = orc.make_camera(**c)
    sc = orc.Scene.from_desc(desc)
    p = orc.make_params(c["width"], c["height"], spp, desc.max_bounces, desc.background, desc.seed, sample_offset=sample_offset,
                        spp_total=spp, n_threads=n_threads)
    t0 = time.perf_counter()
    lin, u8, st, _ = sc.render(cam, p)
    wall = time.perf_counter() - t0
    counts = {"rays": int(st.rays), "paths": int(st.paths), "node_visits": int(st.node_visits), "sphere_tests": int(st.sphere_tests),
              "quad_tests": int(st.quad_tests), "hits": int(st.hits)}
    return lin, u8, counts, wall, int(st.threads_used)


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (oracle port: the Rust crate cannot be built here,
    SURVEY.md F8) on all host cores, on the SAME workload; every step renders a bounded sample of it (--ref-spp samples per
    pixel at the full resolution and depth; Mrays/s is a rate, so it does not depend on how many samples are averaged)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    desc = workload_desc(args.workload, args.spp or None)
    spp = min(args.ref_spp, desc.spp)
    warm = max(args.warmup, 0)
    walls, counts, cores = [], None, 0
    for i in range(warm + args.steps):
        _, _, counts, wall, cores = oracle_render_sample(desc, spp, 0, sample_offset=0)
        if i >= warm:
            walls.append(wall)
    ms = 1e3 * float(np.mean(walls))
    mrays = counts["rays"] / (ms / 1e3) / 1e6
    c = desc.camera
    sample = "every step renders %d of the %d samples per pixel (sample indices 0..%d) at the full %dx%d, depth %d" % (spp, desc.spp, spp - 1, c["width"], c["height"], desc.max_bounces)
    out = {"impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "mpaths_per_s": counts["paths"] / (ms / 1e3) / 1e6,
           "config": {"workload": workload_string(desc), "sample": sample,
                      "kernel": "CPU oracle port of the reference renderer (C++17, -O2, no FMA), %d host threads" % cores},
           "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


def image_parity(glin, gu8, olin, ou8):
    mse = float(np.mean((gu8.astype(np.float64) - ou8.astype(np.float64)) ** 2))
    fin = np.isfinite(olin)
    ulp = np.abs(glin[fin].view(np.int32).astype(np.int64) - olin[fin].view(np.int32).astype(np.int64))
    return {"psnr_db_8bit": None if mse == 0 else 10 * np.log10(255.0 ** 2 / mse), "psnr_is_infinite": mse == 0,
            "linear_max_ulp": int(ulp.max(initial=0)), "linear_values_differing": int((ulp > 0).sum()),
            "rgb8_max_lsb": int(np.abs(gu8.astype(np.int32) - ou8.astype(np.int32)).max(initial=0))}


def scene_fits_shared_memory(desc):
    """Mirror of the launcher's rule (trt_kernels.cu scene_smem_bytes <= 64 KB): such scenes are served from shared memory /
    the constant bank and are bound by instruction issue; larger ones by the memory system."""
    ns = int((desc.prim_type == 0).sum()); nq = int((desc.prim_type == 1).sum())
    return 16 * (4 * max(desc.n_prims - 1, 0) + ns + 4 * nq) <= 64 * 1024


def rooflines(args, desc, per_ray, per_ray_src, walked, rays_per_launch, kern_avg_ms, clk, kernel_name):
    """`roofline` = the bound that binds this workload: fp32 issue for shared-memory-resident scenes, the
    memory system (bytes of the tree actually walked) for big ones. The other one is reported beside it, demoted."""
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    traffic_all = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic_all = json.load(open(tp))
    tr = traffic_all.get(args.workload)
    traffic = tr.get("dram_bytes") if isinstance(tr, dict) else tr
    secs = kern_avg_ms / 1e3
    issue = hbm_ref = hbm_walk = None
    if per_ray:
        peak_ops = 148 * 128 * sm_max_mhz * 1e6 / 1e12
        tops = per_ray["ops"] * rays_per_launch / secs / 1e12
        issue = {"bound": "fp32_issue", "achieved": tops, "peak": peak_ops, "unit": "Tlaneop/s", "frac": tops / peak_ops, "traffic": traffic,
                 "kernel": kernel_name, "kernel_ms": kern_avg_ms, "peak_source": "148 SMs x 128 fp32 lanes x %.0f MHz (sm_max_mhz, %s)" % (sm_max_mhz, peak_src),
                 "peak_at_observed_clock": 148 * 128 * (clk.get("sm_mhz") or sm_max_mhz) * 1e6 / 1e12,
                 "algorithmic_ops_per_ray": per_ray["ops"], "events_per_ray": {k: per_ray[k] for k in ("V", "Ps", "Pq", "H")},
                 "events_source": per_ray_src, "ops_per_event": {k: ALGO[k] for k in ("node_ops", "sphere_ops", "quad_ops", "hit_ops")}}
        gbs = per_ray["bytes"] * rays_per_launch / secs / 1e9
        hbm_ref = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
                   "algorithmic_bytes_per_ray": per_ray["bytes"], "peak_source": peak_src,
                   "note": "SURVEY §8d bytes counted on the REFERENCE-topology tree; for a scene served from shared memory / the constant "
                           "bank this is on-chip traffic, not an HBM bound (real DRAM bytes per launch: `traffic`)"}
    if walked:
        gbs = walked["bytes"] * rays_per_launch / secs / 1e9
        hbm_walk = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
                    "kernel": kernel_name, "kernel_ms": kern_avg_ms, "peak_source": peak_src,
                    "algorithmic_bytes_per_ray": walked["bytes"], "records_per_ray": {k: walked[k] for k in ("wide_nodes", "spheres", "quads", "hits")},
                    "bytes_per_record": {"wide_node": 112, "sphere": 16, "quad": 64, "hit": 32, "path": 16},
                    "events_source": walked["source"],
                    "l2_bytes_per_launch": tr.get("l2_bytes") if isinstance(tr, dict) else None,
                    "note": "bytes of the records the walk really fetches (TRT_FLAG_COUNT_WALKED) x rays / kernel time; the scene "
                            "(76 MB at 10^6 spheres) sits in the 126 MB L2, so `traffic` (DRAM) is far below the algorithmic bytes: "
                            "the walk is bound by the LATENCY of dependent L2 fetches, which this fraction of the HBM roofline quantifies"}
    if scene_fits_shared_memory(desc) or hbm_walk is None:
        return issue, {"roofline_hbm": hbm_ref, "roofline_walked": hbm_walk}
    return hbm_walk, {"roofline_issue": issue, "roofline_hbm_reference_tree": hbm_ref}


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c1s", "c2", "c3", "c4", "c5"])
    ap.add_argument("--spp", type=int, default=0, help="override the workload's samples per pixel (0 = BASELINE value)")
    ap.add_argument("--bvh", default="auto", choices=["auto", "sah", "reference", "flat", "lbvh"])
    ap.add_argument("--variant", default="megakernel", choices=["megakernel", "wavefront"])
    ap.add_argument("--ref-spp", type=int, default=4, help="--impl reference: samples per pixel per step (bounded sample)")
    ap.add_argument("--cpu-spp", type=int, default=8, help="cpu_baseline / parity leg: samples per pixel of the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle legs (cpu_baseline and parity)")
    ap.add_argument("--native-multi", action="store_true",
                    help="one process, --gpus N devices through trt_multi_render (threads + NVLink peer reduce inside the C ABI) instead of one "
                         "torch.distributed rank per GPU; not for use under torchrun")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.native_multi:
        return run_native_multi(args)

    import torch
    import torch.distributed as dist
    import tiny_raytracer_b200 as trt
    from tiny_raytracer_b200 import multigpu

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    desc = workload_desc(args.workload, args.spp or None)
    c = desc.camera; W, H, SPP = c["width"], c["height"], desc.spp
    kind = {"auto": trt.BVH_AUTO, "sah": trt.BVH_SAH, "reference": trt.BVH_REFERENCE, "flat": trt.BVH_FLAT, "lbvh": trt.BVH_LBVH_DEVICE}[args.bvh]
    variant = trt.VARIANT_MEGAKERNEL if args.variant == "megakernel" else trt.VARIANT_WAVEFRONT
    flat = args.bvh == "flat" or (args.bvh == "auto" and 1 <= desc.n_prims <= 64)          # TRT_BVH_AUTO's rule (trt_scene.cpp)
    bvh_resolved = "flat" if flat else (("lbvh" if desc.n_prims >= 65536 else "sah") if args.bvh == "auto" else args.bvh)
    kernel_name = "k_wf_shade + k_wf_extend" if variant != 0 else ("k_render_mega_flat" if flat else "k_render_mega")
    ctx = trt.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)   # same stream as torch's events and the NCCL reduce
    cam = trt.make_camera(**c)
    scene = trt.Scene.from_desc(desc, kind)
    scene.upload(ctx)

    def mk(spp, off, tot, flags=0):
        return trt.make_params(W, H, spp, desc.max_bounces, desc.background, desc.seed, sample_offset=off, spp_total=tot, variant=variant, flags=flags)

    n_val = H * W * 3
    acc = torch.zeros(n_val, dtype=torch.float64, device=dev)
    lin = torch.empty(n_val, dtype=torch.float32, device=dev)
    u8 = torch.empty(n_val, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step(kernel_events=None, want_stats=False, spp_total=SPP, flags=0):
        acc.zero_()
        if kernel_events:
            kernel_events[0].record()
        st = multigpu.render_share(ctx, scene, cam, lambda s, o, t: mk(s, o, t, flags), spp_total, world, rank, acc, want_stats=want_stats)
        if kernel_events:
            kernel_events[1].record()
        multigpu.reduce_accumulators(acc, 0)
        if rank == 0:
            trt.finalize(ctx, acc.data_ptr(), W, H, spp_total, lin.data_ptr(), u8.data_ptr())
        return st

    def all_sum(values):
        t = torch.tensor(values, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    # ---- warm-up (>= 3), the first one also reads the ray / path counts of this rank's share
    st0 = step(want_stats=True)
    my_rays, my_paths = (int(st0.rays), int(st0.paths)) if st0 is not None else (0, 0)
    render_launches = int(st0.launches) if st0 is not None and int(st0.launches) > 0 else 1   # 2 where the first-bounce pass runs (k_primary + path loop)
    for _ in range(max(args.warmup, 3) - 1):
        step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- timed region: K steps, per-step CUDA events on the launching stream, L2 flushed between steps (not timed)
    clocks = ClockSampler(local); clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier(); torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step(kernel_events=kev[k])
        ev[k][1].record()
    torch.cuda.synchronize(); barrier()
    wall_timed = time.perf_counter() - wall0
    clk = clocks.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    t = torch.tensor([sum(step_ms), float(np.mean(kern_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_avg_ms = float(t[0]), float(t[1])
    rays, paths = all_sum([float(my_rays), float(my_paths)])
    ms_per_step = total_ms / args.steps
    mrays = rays / (ms_per_step / 1e3) / 1e6
    mpaths = paths / (ms_per_step / 1e3) / 1e6

    # ---- e2e: the call a user of the reference makes, HOST buffers, copies inside the timed region (wall clock, max over ranks).
    #      N = 1: literally trt_render (Renderer::render's replacement): re-commit (host BVH build) + upload (H2D) + render +
    #             tone map + D2H of the f32 and the 8-bit image into pinned host memory, all inside that one call.
    #      N > 1: the same sequence per rank with the NCCL reduce between render and tone map (rank 0 copies the image home).
    host_lin = torch.empty(n_val, dtype=torch.float32).pin_memory() if rank == 0 else None
    host_u8 = torch.empty(n_val, dtype=torch.uint8).pin_memory() if rank == 0 else None
    full = mk(SPP, 0, SPP)

    def e2e_step():
        scene.commit(kind)            # host BVH build: bumps the version so that the upload really copies
        if world == 1:
            scene.render(ctx, cam, full, out_linear=host_lin.numpy(), out_rgb8=host_u8.numpy())      # trt_render
        else:
            scene.upload(ctx)
            step()
            if rank == 0:
                host_lin.copy_(lin, non_blocking=True); host_u8.copy_(u8, non_blocking=True)
            torch.cuda.synchronize()

    e2e_step()
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    info = scene.bvh_info()
    scene_bytes = 64 * info["nodes"] + 16 * int((desc.prim_type == 0).sum()) + 64 * int((desc.prim_type == 1).sum()) + 16 * desc.n_prims + 32 * len(desc.materials)
    e2e_mrays = rays / e2e_s / 1e6

    # ---- the records the walk really fetches (tree kinds): one counted render of a small sample, every rank its share
    walked = None
    if not flat and variant == 0:
        cspp = max(min(2, SPP), 1) if desc.n_prims > 100_000 else min(8, SPP)
        cspp = max(cspp, min(world, SPP))
        stc = step(want_stats=True, spp_total=cspp, flags=trt._lib.FLAG_COUNT_WALKED)
        v = [float(getattr(stc, k)) if stc is not None else 0.0 for k in ("rays", "paths", "node_visits", "sphere_tests", "quad_tests")]
        cr, cp, cv, cs, cq = all_sum(v)
        walked = {"rays": cr, "paths": cp, "wide_nodes": cv / max(cr, 1), "spheres": cs / max(cr, 1), "quads": cq / max(cr, 1),
                  "source": "TRT_FLAG_COUNT_WALKED render of %d spp, this run (%s)" % (cspp, "4-wide records" if info["nodes"] > 1 else "binary records")}

    # ---- parity at EVERY N (VERDICT r1 #1a): every rank renders its share of --cpu-spp samples, reduce, tone map, and rank 0
    #      compares the image with the oracle's render of the same sample indices; ray counts are summed over the ranks
    do_oracle = not args.no_cpu_baseline
    cpu_spp = min(max(args.cpu_spp, min(world, SPP)), SPP)
    par_rays = None
    if do_oracle:
        stp = step(want_stats=True, spp_total=cpu_spp)
        torch.cuda.synchronize()
        par_rays = all_sum([float(stp.rays) if stp is not None else 0.0])[0]
        if rank == 0:
            glin = lin.cpu().numpy().reshape(H, W, 3); gu8 = u8.cpu().numpy().reshape(H, W, 3)
    if rank != 0:
        if world > 1:
            dist.barrier()        # rank 0 is running the oracle
            dist.destroy_process_group()
        return 0

    cpu_baseline, parity, per_ray, per_ray_src = None, None, None, None
    if do_oracle:
        olin, ou8, counts, wall, cores = oracle_render_sample(desc, cpu_spp)
        per_ray = per_ray_from_counts(counts); per_ray_src = "oracle, this run (%d spp sample, reference-topology tree)" % cpu_spp
        if world == 1:           # the contract's cpu_baseline leg: rank 0 at N = 1 only
            cpu_baseline = {"value": counts["rays"] / wall / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                            "sample": "%dx%d, %d spp (sample indices 0..%d), depth %d; %.1f s wall" % (W, H, cpu_spp, cpu_spp - 1, desc.max_bounces, wall),
                            "mpaths_per_s": counts["paths"] / wall / 1e6}
        parity = image_parity(glin, gu8, olin, ou8)
        parity.update({"rays_equal": int(par_rays) == counts["rays"], "n_gpus": world,
                       "sample": "%d spp, same Philox sample indices on both sides, partitioned over %d GPU(s) and reduced" % (cpu_spp, world)})
        if world == 1:
            # primary hits (the (prim_id, t) mismatch count must be 0): one sample per pixel on both sides
            from oracle import pyoracle as orc
            op1 = orc.make_params(W, H, 1, desc.max_bounces, desc.background, desc.seed)
            oprim, ot, _, _ = orc.Scene.from_desc(desc).render_aov(orc.make_camera(**c), op1)
            gprim, gt, _, _ = scene.render_aov(ctx, cam, mk(1, 0, 1))
            parity["primary_hit_mismatches"] = int(((gprim != oprim) | (gt.view(np.uint32) != ot.view(np.uint32))).sum())
            parity["primary_hits_compared"] = int(oprim.size)
    if per_ray is None:
        g = golden_counters(args.workload)
        if g:
            per_ray = per_ray_from_counts(g); per_ray_src = "tests/golden/workload_counters.json (oracle, reference-topology tree)"
    if walked is not None:
        hits = per_ray["H"] if per_ray else (walked["rays"] - walked["paths"]) / max(walked["rays"], 1)
        walked["hits"] = hits
        walked["bytes"] = 112 * walked["wide_nodes"] + 16 * walked["spheres"] + 64 * walked["quads"] + 32 * hits + 16 * walked["paths"] / max(walked["rays"], 1)
    if render_launches == 2 and variant == 0:
        kernel_name = "k_primary + " + kernel_name       # the coherent first-bounce pass runs before the path loop; kernel_ms covers both
    roofline, other = rooflines(args, desc, per_ray, per_ray_src, walked, rays / world, kern_avg_ms, clk, kernel_name)

    out = {"metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "mpaths_per_s": mpaths, "rays_per_step": rays, "paths_per_step": paths,
           "config": {"workload": workload_string(desc), "bvh": args.bvh, "bvh_resolved": bvh_resolved, "kernel": kernel_name, "variant": args.variant,
                      "parallelism": "spp-partitioned x%d + NCCL reduce(f64 sum)" % world if world > 1 else "single GPU",
                      "l2": "256 MiB memset between timed steps (outside the events)", "timing": "per-step CUDA events on the launch stream, max over ranks"},
           "clocks": clk,
           "e2e": {"value": e2e_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes + 256), "d2h_bytes_per_step": int(n_val * 5),
                   "ms_per_step": e2e_s * 1e3,
                   "what": ("trt_render (the C-ABI call behind Renderer::render): host BVH commit + scene upload + render + tone map + f32 and 8-bit image into pinned host buffers"
                            if world == 1 else "per rank: commit + upload scene, render share, NCCL reduce, tone map, rank 0 copies f32 + 8-bit image to pinned host memory")},
           "gpu_launches": int(args.steps * (world * render_launches + 1)),   # per step: the render kernel(s) of every rank + one k_finalize on rank 0
           "wall_s_timed_region": wall_timed,
           "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity}
    out.update(other)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_native_multi(args):
    """One process, N GPUs through the C ABI (trt_multi_render): what a C / C++ / Rust host gets from ONE call. `value` is
    device-timed inside the library (max over the devices of the render kernel + max of the reduce/tone-map kernel, CUDA
    events on each device's stream); `e2e` is the wall clock of the whole call with host buffers."""
    import torch
    import tiny_raytracer_b200 as trt
    n = args.gpus
    desc = workload_desc(args.workload, args.spp or None)
    c = desc.camera; W, H, SPP = c["width"], c["height"], desc.spp
    kind = {"auto": trt.BVH_AUTO, "sah": trt.BVH_SAH, "reference": trt.BVH_REFERENCE, "flat": trt.BVH_FLAT, "lbvh": trt.BVH_LBVH_DEVICE}[args.bvh]
    cam = trt.make_camera(**c)
    scene = trt.Scene.from_desc(desc, kind)
    mc = trt.MultiContext(list(range(n)))
    p = trt.make_params(W, H, SPP, desc.max_bounces, desc.background, desc.seed, spp_total=SPP)
    n_val = H * W * 3
    host_lin = torch.empty(n_val, dtype=torch.float32).pin_memory(); host_u8 = torch.empty(n_val, dtype=torch.uint8).pin_memory()
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
    for _ in range(max(args.warmup, 3)):
        _, _, st = mc.render(scene, cam, p, want_linear=False, want_rgb8=False)
    clocks = ClockSampler(0); clocks.start()
    dev_ms, red_ms = [], []
    for _ in range(args.steps):
        for i, f in enumerate(flush):
            f.zero_(); torch.cuda.synchronize(i)
        _, _, st = mc.render(scene, cam, p, want_linear=False, want_rgb8=False)     # result stays on the devices
        dev_ms.append(st.device_ms); red_ms.append(st.finalize_ms)
    clk = clocks.stop()
    ms_per_step = float(np.mean(dev_ms) + np.mean(red_ms))
    rays, paths = int(st.rays), int(st.paths)
    mc.render(scene, cam, p, out_linear=host_lin.numpy(), out_rgb8=host_u8.numpy())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        scene.commit(kind)
        glin, gu8, st = mc.render(scene, cam, p, out_linear=host_lin.numpy(), out_rgb8=host_u8.numpy())
    e2e_s = (time.perf_counter() - t0) / args.steps
    parity = None
    if not args.no_cpu_baseline:
        cpu_spp = min(max(args.cpu_spp, n), SPP)
        pp = trt.make_params(W, H, cpu_spp, desc.max_bounces, desc.background, desc.seed, spp_total=cpu_spp)
        glin, gu8, stp = mc.render(scene, cam, pp)
        olin, ou8, counts, wall, cores = oracle_render_sample(desc, cpu_spp)
        parity = image_parity(glin, gu8, olin, ou8)
        parity.update({"rays_equal": int(stp.rays) == counts["rays"], "n_gpus": n, "sample": "%d spp, same Philox sample indices, partitioned over %d GPU(s) inside trt_multi_render" % (cpu_spp, n)})
    info = scene.bvh_info()
    scene_bytes = 64 * info["nodes"] + 16 * int((desc.prim_type == 0).sum()) + 64 * int((desc.prim_type == 1).sum()) + 16 * desc.n_prims + 32 * len(desc.materials)
    out = {"metric": "Mrays/s", "value": rays / (ms_per_step / 1e3) / 1e6, "unit": "Mrays/s", "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "mpaths_per_s": paths / (ms_per_step / 1e3) / 1e6, "rays_per_step": rays, "paths_per_step": paths,
           "render_ms_max_over_devices": float(np.mean(dev_ms)), "reduce_tonemap_ms_max_over_devices": float(np.mean(red_ms)),
           "config": {"workload": workload_string(desc), "bvh": args.bvh, "variant": "megakernel",
                      "parallelism": "native: trt_multi_render, %d devices in one process, spp-partitioned, NVLink peer reduce-scatter + tone map (%s)" % (n, "peer loads" if mc.peer_direct else "staged copies"),
                      "l2": "256 MiB memset per device between timed steps", "timing": "CUDA events inside the library on each device's stream, max over devices"},
           "clocks": clk,
           "e2e": {"value": rays / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(n * (scene_bytes + 256)), "d2h_bytes_per_step": int(n_val * 5),
                   "ms_per_step": e2e_s * 1e3, "what": "host BVH commit + trt_multi_render: upload to every device, render shares, peer reduce-scatter + tone map, every device copies its slice into the pinned host image"},
           "gpu_launches": int(args.steps * 2 * n), "parity": parity}
    print(json.dumps(out), flush=True)
    mc.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
