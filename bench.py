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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    desc = workload_desc(args.workload)
    spp = args.ref_spp
    walls, counts, cores = [], None, 0
    for i in range(args.warmup_ref + args.steps):
        _, _, counts, wall, cores = oracle_render_sample(desc, spp, 0, sample_offset=0)
        if i >= args.warmup_ref:
            walls.append(wall)
    ms = 1e3 * float(np.mean(walls))
    mrays = counts["rays"] / (ms / 1e3) / 1e6
    c = desc.camera
    sample = "%dx%d, %d spp (sample indices 0..%d), depth %d per step" % (c["width"], c["height"], spp, spp - 1, desc.max_bounces)
    out = {"impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup_ref, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "mpaths_per_s": counts["paths"] / (ms / 1e3) / 1e6,
           "config": {"workload": workload_string(desc), "kernel": "CPU oracle port of the reference renderer (C++17, -O2, no FMA)"},
           "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


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
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--cpu-spp", type=int, default=8, help="cpu_baseline leg: samples per pixel of the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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

    def step(kernel_events=None, want_stats=False):
        acc.zero_()
        if kernel_events:
            kernel_events[0].record()
        st = multigpu.render_share(ctx, scene, cam, mk, SPP, world, rank, acc, want_stats=want_stats)
        if kernel_events:
            kernel_events[1].record()
        multigpu.reduce_accumulators(acc, 0)
        if rank == 0:
            trt.finalize(ctx, acc.data_ptr(), W, H, SPP, lin.data_ptr(), u8.data_ptr())
        return st

    # ---- warm-up (>= 3), the first one also reads the ray / path counts of this rank's share
    st0 = step(want_stats=True)
    my_rays, my_paths = (int(st0.rays), int(st0.paths)) if st0 is not None else (0, 0)
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
    t = torch.tensor([sum(step_ms), float(np.mean(kern_ms)), float(my_rays), float(my_paths)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, kern_avg_ms, rays, paths = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tsum[3])
    else:
        total_ms, kern_avg_ms, rays, paths = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ms_per_step = total_ms / args.steps
    mrays = rays / (ms_per_step / 1e3) / 1e6
    mpaths = paths / (ms_per_step / 1e3) / 1e6

    # ---- e2e: the public host-buffer path. Every step re-commits and re-uploads the scene (H2D), renders, reduces,
    #      tone-maps and copies the linear + 8-bit image back to pinned host memory (D2H). Wall clock, max over ranks.
    scene_bytes = None
    host_lin = torch.empty(n_val, dtype=torch.float32).pin_memory() if rank == 0 else None
    host_u8 = torch.empty(n_val, dtype=torch.uint8).pin_memory() if rank == 0 else None

    def e2e_step():
        scene.commit(kind)            # host BVH build: bumps the version so that upload really copies
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- cpu_baseline + parity (rank 0, N = 1 only): the oracle on a bounded sample of the same workload
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    cpu_baseline, parity, per_ray, per_ray_src = None, None, None, None
    if world == 1 and not args.no_cpu_baseline:
        cpu_spp = min(args.cpu_spp, SPP)
        olin, ou8, counts, wall, cores = oracle_render_sample(desc, cpu_spp)
        per_ray = per_ray_from_counts(counts); per_ray_src = "oracle, this run (%d spp sample, reference-topology tree)" % cpu_spp
        cpu_baseline = {"value": counts["rays"] / wall / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                        "sample": "%dx%d, %d spp (sample indices 0..%d), depth %d; %.1f s wall" % (W, H, cpu_spp, cpu_spp - 1, desc.max_bounces, wall),
                        "mpaths_per_s": counts["paths"] / wall / 1e6}
        glin, gu8, gst = scene.render(ctx, cam, mk(cpu_spp, 0, cpu_spp))
        mse = float(np.mean((gu8.astype(np.float64) - ou8.astype(np.float64)) ** 2))
        fin = np.isfinite(olin)
        ulp = np.abs(glin[fin].view(np.int32).astype(np.int64) - olin[fin].view(np.int32).astype(np.int64))
        parity = {"psnr_db_8bit": None if mse == 0 else 10 * np.log10(255.0 ** 2 / mse), "psnr_is_infinite": mse == 0,
                  "linear_max_ulp": int(ulp.max(initial=0)), "linear_values_differing": int((ulp > 0).sum()),
                  "rays_equal": int(gst.rays) == counts["rays"], "sample": "%d spp, same Philox sample indices on both sides" % cpu_spp}
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
    roofline, roofline_issue = None, None
    if per_ray:
        rays_per_launch = rays / world     # the slowest rank's launch carries ~1/N of the rays
        gbs = per_ray["bytes"] * rays_per_launch / (kern_avg_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        roofline = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
                    "kernel": kernel_name, "kernel_ms": kern_avg_ms, "peak_source": peak_src,
                    "algorithmic_bytes_per_ray": per_ray["bytes"], "events_per_ray": {k: per_ray[k] for k in ("V", "Ps", "Pq", "H")},
                    "events_source": per_ray_src,
                    "note": "scene is cache-resident for Cornell-sized workloads: the binding bound is fp32 issue, see roofline_issue"}
        clk_mhz = clk.get("sm_mhz") or sm_max_mhz
        peak_ops = 148 * 128 * sm_max_mhz * 1e6 / 1e12
        tops = per_ray["ops"] * rays_per_launch / (kern_avg_ms / 1e3) / 1e12
        roofline_issue = {"bound": "fp32_issue", "achieved": tops, "peak": peak_ops, "unit": "Tlaneop/s", "frac": tops / peak_ops,
                          "peak_at_observed_clock": 148 * 128 * clk_mhz * 1e6 / 1e12, "algorithmic_ops_per_ray": per_ray["ops"]}

    out = {"metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "mpaths_per_s": mpaths, "rays_per_step": rays, "paths_per_step": paths,
           "config": {"workload": workload_string(desc), "bvh": args.bvh, "bvh_resolved": bvh_resolved, "kernel": kernel_name, "variant": args.variant,
                      "parallelism": "spp-partitioned x%d + NCCL reduce(f64 sum)" % world if world > 1 else "single GPU",
                      "l2": "256 MiB memset between timed steps (outside the events)", "timing": "per-step CUDA events on the launch stream, max over ranks"},
           "clocks": clk,
           "e2e": {"value": e2e_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes + 256), "d2h_bytes_per_step": int(n_val * 5),
                   "ms_per_step": e2e_s * 1e3, "what": "commit + upload scene, render share, NCCL reduce, tone-map, copy f32+u8 image to pinned host"},
           "gpu_launches": int(args.steps * (world + 1)),   # per step: one k_render_mega per rank + one k_finalize on rank 0
           "wall_s_timed_region": wall_timed,
           "roofline": roofline, "roofline_issue": roofline_issue, "cpu_baseline": cpu_baseline, "parity": parity}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
