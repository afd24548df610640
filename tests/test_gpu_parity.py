"""GPU parity tests: the CUDA path (through the C ABI, libtrt_b200.so) against the CPU oracle on the same seeded inputs.

Bar (north_star): primary-hit primitive ids and t BIT-EXACT; per-sample radiance bit-exact (same IEEE op sequence);
linear image bit-exact when a pixel's samples are summed by one lane, <= 1 ulp otherwise; 8-bit image within 1 LSB
(device powf vs glibc powf) and PSNR >= 40 dB.
"""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PSNR_MIN_DB = 40.0   # north_star tolerance


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def both(trt, orc, desc):
    ocam = orc.make_camera(**desc.camera)
    cam = trt.make_camera(**desc.camera)
    return orc.Scene.from_desc(desc), ocam, cam


def params(trt, orc, desc, **kw):
    w, h = desc.camera["width"], desc.camera["height"]
    spp = kw.pop("spp", desc.spp); mb = kw.pop("max_bounces", desc.max_bounces)
    off = kw.pop("sample_offset", 0); tot = kw.pop("spp_total", 0); flags = kw.pop("flags", 0)
    return (trt.make_params(w, h, spp, mb, desc.background, desc.seed, sample_offset=off, spp_total=tot, flags=flags),
            orc.make_params(w, h, spp, mb, desc.background, desc.seed, sample_offset=off, spp_total=tot))


KINDS = [0, 1, 2]   # TRT_BVH_REFERENCE, TRT_BVH_SAH, TRT_BVH_FLAT (scenes of 1..64 primitives only: TRT_FLAT_MAX)


def flat_ok(kind, desc):
    if kind == 2 and not (1 <= desc.n_prims <= 64):
        pytest.skip("TRT_BVH_FLAT takes 1..64 primitives")


# ------------------------------------------------------------------ unit level: the numeric contract on the device
def test_philox_device_matches_oracle_and_kat(trt, orc, ctx):
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        got = trt.debug_philox(ctx, np.array([ctr], np.uint32), key)[0]
        assert [int(x) for x in got] == want
    rng = np.random.default_rng(0)
    ctrs = rng.integers(0, 2 ** 32, (4096, 4), dtype=np.uint64).astype(np.uint32)
    key = [123456789, 987654321]
    got = trt.debug_philox(ctx, ctrs, key)
    want = np.array([orc.philox(c, key) for c in ctrs[:512]], np.uint32)
    assert np.array_equal(got[:512], want)


def test_det_math_device_bit_exact(trt, orc, ctx):
    u = (np.arange(0, 1 << 23, 37, dtype=np.uint32).astype(np.float32)) * np.float32(2.0 ** -23)
    u = np.concatenate([u, np.float32([0.0, 2.0 ** -23, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1.0 - 2.0 ** -23])])
    s, c, cb = trt.debug_math(ctx, u)
    os_, oc, ocb = orc.det_math(u)
    assert np.array_equal(bits(s), bits(os_)) and np.array_equal(bits(c), bits(oc)) and np.array_equal(bits(cb), bits(ocb))


# ------------------------------------------------------------------ closest hit on arbitrary rays
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("scene", ["cornell", "cornell_spheres", "five_spheres", "five_quads", "mixed200", "mixed800", "final"])
def test_trace_rays_bit_exact(trt, orc, ctx, kind, scene):
    S = trt.scenes
    desc = {"cornell": S.cornell, "cornell_spheres": S.cornell_spheres, "five_spheres": S.five_spheres, "five_quads": S.five_quads,
            "mixed200": lambda: S.random_mixed(200, 3), "mixed800": lambda: S.random_mixed(800, 4), "final": S.final_scene}[scene]()
    flat_ok(kind, desc)
    osc = orc.Scene.from_desc(desc)
    sc = trt.Scene.from_desc(desc, kind)
    rng = np.random.default_rng(42)
    n = 200_000
    lo = desc.prim_data[:, :3].min(0) - 3; hi = desc.prim_data[:, :3].max(0) + 3
    lo, hi = np.maximum(lo, -120), np.minimum(hi, 120)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # degenerate directions: exact zeros, negative zeros, axis-aligned; origins exactly on planes
    d[:2000, 0] = 0.0; d[2000:4000, 1] = -0.0; d[4000:5000] = np.float32([0, 0, 1]); d[5000:6000] = np.float32([0, -1, 0])
    o[6000:8000] = np.round(o[6000:8000])
    prim, t = sc.trace_rays(ctx, o, d, 0.001)
    oprim, ot = osc.trace_rays(o, d, 0.001, use_bvh=True)
    assert np.array_equal(prim, oprim), "%d prim mismatches" % int((prim != oprim).sum())
    assert np.array_equal(bits(t), bits(ot))
    assert (prim != 0xFFFFFFFF).sum() > n // 50


# ------------------------------------------------------------------ primary-hit AOV + whole-path radiance
AOV_CASES = {
    "c1": lambda S: S.cornell(400, 400, 16, 20),                       # BASELINE config 1, full size
    "c1s": lambda S: S.cornell_spheres(400, 400, 8, 20),
    "five_spheres": lambda S: S.five_spheres(400, 300, 3, 10),         # renderer.rs:125-150
    "five_quads": lambda S: S.five_quads(400, 300, 10, 10),            # quad.rs:98-151
    "final_small": lambda S: S.final_scene(300, 169, 4, 50),
    "glass_small": lambda S: S.glass_cornell(240, 135, 8, 100),
    "mixed": lambda S: S.random_mixed(400, 21, 160, 120, 4, 16),
    "one_pixel": lambda S: S.cornell(1, 1, 5, 20),                     # W-1 = 0 divisor: inf/NaN rays must agree too
    "ragged": lambda S: S.cornell_spheres(37, 19, 3, 7),               # not a multiple of the 8x4 tile
    "single_prim": lambda S: S.random_mixed(1, 5, 32, 32, 2, 8),
}


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", sorted(AOV_CASES))
def test_aov_and_radiance_bit_exact(trt, orc, ctx, kind, case):
    desc = AOV_CASES[case](trt.scenes)
    flat_ok(kind, desc)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, kind)
    p, op = params(trt, orc, desc)
    prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
    oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True)
    assert np.array_equal(prim, oprim), "%d primary-hit id mismatches" % int((prim != oprim).sum())
    assert np.array_equal(bits(t), bits(ot)), "primary-hit t differs"
    assert np.array_equal(nr, onr), "path length differs on %d samples" % int((nr != onr).sum())
    assert np.array_equal(bits(rad), bits(orad)), "radiance differs on %d values" % int((bits(rad) != bits(orad)).sum())


@pytest.mark.parametrize("kind", KINDS)
def test_primary_hits_full_resolution(trt, orc, ctx, kind):
    """BASELINE configs 2, 3 and 5 at their full resolution, 1 sample per pixel, primary hit only."""
    S = trt.scenes
    for desc in (S.final_scene(1200, 675, 1, 50), S.cornell(1920, 1080, 1, 50), S.glass_cornell(3840, 2160, 1, 100)):
        if kind == 2 and desc.n_prims > 64:
            continue
        osc, ocam, cam = both(trt, orc, desc)
        sc = trt.Scene.from_desc(desc, kind)
        p, op = params(trt, orc, desc)
        prim, t, _, _ = sc.render_aov(ctx, cam, p)
        oprim, ot, _, _ = osc.render_aov(ocam, op)
        assert np.array_equal(prim, oprim), "%s: %d mismatches" % (desc.name, int((prim != oprim).sum()))
        assert np.array_equal(bits(t), bits(ot))


@pytest.mark.parametrize("kind", KINDS)
def test_primary_hits_bvh_stress(trt, orc, ctx, kind):
    """BASELINE config 4 scene family (random spheres) at 200k spheres, 960x540, primary + two bounces.

    TRT_BVH_REFERENCE reproduces the reference by construction: everything bit-exact.
    TRT_BVH_SAH visits nodes in another order. The reference's result is order-dependent in one corner:
    a grazing sphere hit whose computed t is smaller than its own box's computed entry (fp cancellation in
    hb*hb - a*c at distance ~150) is culled or not depending on which primitive the reference happened to visit first.
    The ordered walk detects exactly that situation on its winner and replays the ray in the reference's order, so it is
    bit-exact as well (before the replay existed, 3-5 of these 0.5 M samples differed)."""
    desc = trt.scenes.random_spheres(200_000, 960, 540, 1, 3)
    flat_ok(kind, desc)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, kind)
    p, op = params(trt, orc, desc)
    prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
    oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True)
    same = (nr == onr) & (bits(rad) == bits(orad)).all(1) & (prim == oprim) & (bits(t) == bits(ot))
    assert same.all(), "%d mismatches" % int((~same).sum())


def test_sah_order_dependent_rays_are_replayed(trt, orc, ctx):
    """Deep paths through 200 k spheres (depth 50, 1 M samples): the ordered walk with its replay equals the oracle on every
    sample, on the host-built SAH tree and on the device-built tree (whose reference tree is built on the device too)."""
    desc = trt.scenes.random_spheres(200_000, 960, 540, 2, 50)
    osc, ocam, cam = both(trt, orc, desc)
    p, op = params(trt, orc, desc)
    oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True)
    for kind in (trt.BVH_SAH, trt.BVH_LBVH_DEVICE):
        prim, t, rad, nr = trt.Scene.from_desc(desc, kind).render_aov(ctx, cam, p, True, True)
        same = (nr == onr) & (bits(rad) == bits(orad)).all(1) & (prim == oprim) & (bits(t) == bits(ot))
        assert same.all(), "kind %d: %d mismatches" % (kind, int((~same).sum()))


# ------------------------------------------------------------------ whole renders (megakernel)
RENDER_CASES = {
    "c1": lambda S: S.cornell(400, 400, 16, 20),
    "c1s": lambda S: S.cornell_spheres(200, 200, 16, 20),
    "five_spheres": lambda S: S.five_spheres(400, 300, 3, 10),
    "final_small": lambda S: S.final_scene(300, 169, 8, 50),
    "glass_small": lambda S: S.glass_cornell(240, 135, 16, 100),
    "ragged": lambda S: S.cornell_spheres(37, 19, 5, 7),
    "tiny": lambda S: S.cornell(3, 2, 64, 20),
}


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("case", sorted(RENDER_CASES))
def test_render_matches_oracle(trt, orc, ctx, kind, case):
    desc = RENDER_CASES[case](trt.scenes)
    flat_ok(kind, desc)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, kind)
    p, op = params(trt, orc, desc)
    lin, u8, st = sc.render(ctx, cam, p)
    olin, ou8, ost, _ = osc.render(ocam, op)
    w, h = desc.camera["width"], desc.camera["height"]
    assert st.paths == w * h * desc.spp == ost.paths            # pointgen.rs:63-107: exactly spp samples per pixel
    assert st.rays == ost.rays
    finite = np.isfinite(olin)
    assert np.array_equal(np.isfinite(lin), finite)
    d = ulp_diff(lin[finite], olin[finite])
    assert d.max(initial=0) <= 1, "linear image: max %d ulp, %d values differ" % (int(d.max()), int((d > 0).sum()))
    assert np.abs(u8.astype(np.int32) - ou8.astype(np.int32)).max() <= 1
    assert psnr8(u8, ou8) >= PSNR_MIN_DB


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("regen_min", [1, 5, 16, 31])
def test_async_regeneration_threshold_does_not_change_results(trt, orc, ctx, kind, regen_min):
    """TRT_FLAG_REGEN_MIN switches the megakernel to its step-wise form (shade/regenerate as soon as n lanes are waiting),
    the schedule used automatically for big trees. It must not change a single bit of a whole-pixel-item image."""
    extra = 0
    for desc in (trt.scenes.cornell_spheres(96, 64, 6, 20), trt.scenes.final_scene(80, 45, 5, 50)):
        osc, ocam, cam = both(trt, orc, desc)
        sc = trt.Scene.from_desc(desc, kind)
        p, op = params(trt, orc, desc, flags=(regen_min << 8) | extra | 2)
        lin, u8, st = sc.render(ctx, cam, p)
        olin, ou8, ost, _ = osc.render(ocam, op)
        assert (st.paths, st.rays) == (ost.paths, ost.rays)
        assert np.array_equal(bits(lin), bits(olin)) and np.abs(u8.astype(np.int32) - ou8.astype(np.int32)).max() <= 1
        p, op = params(trt, orc, desc, flags=(regen_min << 8) | extra)          # with the tail split: f64 atomics, <= 1 ulp
        lin, u8, st = sc.render(ctx, cam, p)
        assert (st.paths, st.rays) == (ost.paths, ost.rays) and ulp_diff(lin, olin).max() <= 1


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("regen_min", [0, 12])
def test_binary_and_wide_nodes_agree(trt, orc, ctx, kind, regen_min):
    """Tree scenes are walked on 4-wide nodes (the binary tree collapsed by two levels) unless TRT_FLAG_BINARY_NODES asks
    for the binary walk: same boxes, same acceptance rules, so the images are identical bit for bit (whole-pixel items:
    fixed summation order) and both match the oracle."""
    for desc in (trt.scenes.final_scene(120, 68, 6, 50), trt.scenes.random_mixed(300, 5, 96, 64, 4, 12), trt.scenes.random_spheres(6000, 96, 54, 3, 10)):
        osc, ocam, cam = both(trt, orc, desc)
        sc = trt.Scene.from_desc(desc, kind)
        imgs = []
        for extra in (0, 4):   # 4 = TRT_FLAG_BINARY_NODES
            p, op = params(trt, orc, desc, flags=(regen_min << 8) | 2 | extra)
            lin, u8, st = sc.render(ctx, cam, p)
            imgs.append((lin, st))
        assert np.array_equal(bits(imgs[0][0]), bits(imgs[1][0])) and imgs[0][1].rays == imgs[1][1].rays
        olin, ou8, ost, _ = osc.render(ocam, op)
        assert imgs[0][1].rays == ost.rays and ulp_diff(imgs[0][0], olin).max() <= 1


def test_big_tree_render_matches_oracle(trt, orc, ctx):
    """20 k spheres (> 4096 nodes): the launcher picks the step-wise schedule by itself; SAH and reference kinds, both
    equal to the oracle (the SAH walk replays the order-dependent rays of DESIGN.md §4.3 in the step-wise loop too)."""
    desc = trt.scenes.random_spheres(20_000, 160, 90, 4, 12)
    osc, ocam, cam = both(trt, orc, desc)
    p, op = params(trt, orc, desc)
    olin, ou8, ost, _ = osc.render(ocam, op)
    for kind in (0, 1):
        sc = trt.Scene.from_desc(desc, kind)
        for flags in (0, 1, 8):      # plain, and both event-counting builds
            p, _ = params(trt, orc, desc, flags=flags)
            lin, u8, st = sc.render(ctx, cam, p)
            assert st.paths == ost.paths
            bad = (ulp_diff(lin, olin) > 1).any(axis=2)
            assert st.rays == ost.rays and not bad.any()
            assert psnr8(u8, ou8) >= PSNR_MIN_DB


@pytest.mark.parametrize("kind", [0, 1, 3])
@pytest.mark.parametrize("case", ["c1", "c1s", "final_small", "glass_small", "ragged", "tiny"])
def test_wavefront_render_matches_oracle(trt, orc, ctx, kind, case):
    """TRT_VARIANT_WAVEFRONT traces the same paths (same Philox keys) through queues in HBM; only the order of the f64
    accumulation differs, so the image must agree with the oracle to <= 1 ulp and ray / path counts exactly."""
    desc = RENDER_CASES[case](trt.scenes)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, kind)
    w, h = desc.camera["width"], desc.camera["height"]
    p = trt.make_params(w, h, desc.spp, desc.max_bounces, desc.background, desc.seed, variant=trt.VARIANT_WAVEFRONT)
    _, op = params(trt, orc, desc)
    lin, u8, st = sc.render(ctx, cam, p)
    olin, ou8, ost, _ = osc.render(ocam, op)
    assert (st.paths, st.rays) == (ost.paths, ost.rays)
    d = ulp_diff(lin, olin)
    assert d.max() <= 1, "linear image: max %d ulp, %d values differ" % (int(d.max()), int((d > 0).sum()))
    assert np.abs(u8.astype(np.int32) - ou8.astype(np.int32)).max() <= 1 and psnr8(u8, ou8) >= PSNR_MIN_DB
    mega, _, _ = sc.render(ctx, cam, params(trt, orc, desc)[0])
    assert ulp_diff(lin, mega).max() <= 1


@pytest.mark.parametrize("kind", KINDS)
def test_event_counters_match_oracle(trt, orc, ctx, kind):
    """Node visits and primitive tests are topology-dependent: with TRT_BVH_REFERENCE they must equal the oracle's
    counts on the reference tree exactly (same tree, same order, same culling); with SAH only rays/paths must."""
    if kind == 2:
        pytest.skip("no tree, no node events")
    for desc in (trt.scenes.cornell(128, 128, 4, 20), trt.scenes.final_scene(160, 90, 4, 50), trt.scenes.random_mixed(300, 9, 96, 64, 4, 12)):
        osc, ocam, cam = both(trt, orc, desc)
        sc = trt.Scene.from_desc(desc, kind)
        p, op = params(trt, orc, desc, flags=1)
        _, _, st = sc.render(ctx, cam, p)
        _, _, ost, _ = osc.render(ocam, op)
        assert (st.rays, st.paths) == (ost.rays, ost.paths)
        if kind == 0:
            assert (st.node_visits, st.sphere_tests, st.quad_tests) == (ost.node_visits, ost.sphere_tests, ost.quad_tests)
        else:
            assert st.node_visits > 0 and st.sphere_tests + st.quad_tests > 0


@pytest.mark.parametrize("case", ["c3", "c5"])
def test_flat_list_equals_reference_tree_at_full_resolution(trt, orc, ctx, case):
    """The two kinds that reproduce the reference by construction — the flat list with its conservative FFMA2 box phase
    and the reference-topology tree — on the BASELINE frames at full resolution (Cornell 1080p x 8 spp: 74 M ray segments;
    glass Cornell 4K x 2 spp, depth 100): identical ray counts and bit-identical images (whole-pixel items, so the f64
    sums are taken in the same order)."""
    desc = trt.scenes.cornell(1920, 1080, 8, 50) if case == "c3" else trt.scenes.glass_cornell(3840, 2160, 2, 100)
    cam = trt.make_camera(**desc.camera)
    out = []
    for kind in (trt.BVH_FLAT, trt.BVH_REFERENCE):
        p, _ = params(trt, orc, desc, flags=2)      # TRT_FLAG_NO_TAIL_SPLIT
        lin, u8, st = trt.Scene.from_desc(desc, kind).render(ctx, cam, p)
        out.append((lin, u8, st))
    assert out[0][2].rays == out[1][2].rays and out[0][2].paths == out[1][2].paths
    assert np.array_equal(bits(out[0][0]), bits(out[1][0])) and np.array_equal(out[0][1], out[1][1])


def test_render_is_deterministic_and_kind_independent(trt, orc, ctx):
    desc = trt.scenes.cornell_spheres(256, 256, 12, 20)
    cam = trt.make_camera(**desc.camera)
    imgs = []
    for kind in (0, 1, 1, 2, 3):
        sc = trt.Scene.from_desc(desc, kind)
        p, _ = params(trt, orc, desc)
        lin, u8, st = sc.render(ctx, cam, p)
        imgs.append(lin)
    for a, b in zip(imgs, imgs[1:]):
        assert ulp_diff(a, b).max() <= 1


def test_sample_partition_is_exact(trt, orc, ctx):
    """The multi-GPU decomposition on one GPU: rendering sample ranges [0,5), [5,12), [12,16) into one accumulator
    gives the same image as one 16-sample render (f64 sums, <= 1 ulp after rounding to f32), and equals the oracle."""
    import ctypes as C
    import torch
    desc = trt.scenes.cornell(200, 160, 16, 20)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, 1)
    w, h = 200, 160
    acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    for off, n in ((0, 5), (5, 7), (12, 4)):
        p, _ = params(trt, orc, desc, spp=n, sample_offset=off, spp_total=16)
        sc.render_accumulate(ctx, cam, p, acc.data_ptr())
    ctx.synchronize()
    lin_t = torch.empty(h * w * 3, dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    trt.finalize(ctx, acc.data_ptr(), w, h, 16, lin_t.data_ptr(), 0)
    ctx.synchronize()
    lin = lin_t.cpu().numpy().reshape(h, w, 3)
    p, op = params(trt, orc, desc)
    whole, _, _ = sc.render(ctx, cam, p)
    olin, _, _, osum = osc.render(ocam, op, want_sum64=True)
    assert ulp_diff(lin, whole).max() <= 1 and ulp_diff(lin, olin).max() <= 1
    np.testing.assert_allclose(acc.cpu().numpy().reshape(h, w, 3), osum, rtol=1e-12, atol=1e-300)


def test_sample_batch_liveness_and_values(trt, orc, ctx):
    """renderer/sampler/cpu.rs:131-180: 100 rays into the 5-sphere world with CpuSampler::new(1, 2, 0): every sample
    comes back, in order; plus values equal to the oracle's."""
    desc = trt.scenes.liveness_world()
    osc = orc.Scene.from_desc(desc)
    n = 100
    pts = np.zeros(n, trt.SAMPLE_POINT_DTYPE)
    for i in range(n):
        tt = np.float32(i) / np.float32(n)
        x = np.float32(-1.0) * (np.float32(1.0) - tt) + np.float32(1.0) * tt
        pts[i] = (i, 0, (x, 0.0, 0.0), (0.0, 0.0, -1.0))
    for kind in KINDS:
        sc = trt.Scene.from_desc(desc, kind)
        p = trt.make_params(100, 1, 1, 2, (0, 0, 0), desc.seed)
        out = sc.sample_batch(ctx, pts, p)
        assert list(out["x"]) == list(range(n)) and not out["y"].any()
        want = osc.sample_batch(pts, orc.make_params(100, 1, 1, 2, (0, 0, 0), desc.seed))
        assert np.array_equal(bits(out["color"]), bits(want["color"]))
    assert sc.sample_batch(ctx, pts[:0], p).shape[0] == 0          # empty batch


def test_empty_scene_and_zero_work(trt, orc, ctx):
    """Empty world: every path returns the background (cpu.rs:58-61); spp = 0 and max_bounces = 0 render black."""
    desc = trt.scenes.cornell(16, 8, 4, 20)
    empty = trt.scenes.SceneDesc("empty", [], np.zeros(0, np.uint32), np.zeros((0, 12), np.float32), np.zeros(0, np.uint32),
                                 dict(desc.camera), 4, 20, (0.25, 0.5, 0.75), 3)
    cam = trt.make_camera(**empty.camera)
    for kind in (0, 1):
        sc = trt.Scene.from_desc(empty, kind)
        for flags in (0, 9 << 8):        # T-then-S and the step-wise loop
            lin, u8, st = sc.render(ctx, cam, trt.make_params(16, 8, 4, 20, empty.background, 3, flags=flags))
            assert np.array_equal(lin.reshape(-1, 3), np.tile(np.float32([0.25, 0.5, 0.75]), (128, 1))) and st.rays == st.paths == 512
            assert int(u8[0, 0, 1]) == orc.tonemap(0.5)
    sc = trt.Scene.from_desc(desc, 1)
    cam = trt.make_camera(**desc.camera)
    for flags in (0, 9 << 8):
        lin, _, st = sc.render(ctx, cam, trt.make_params(16, 8, 4, 0, desc.background, 1, flags=flags))
        assert not lin.any() and st.rays == 0 and st.paths == 512
        lin, _, st = sc.render(ctx, cam, trt.make_params(16, 8, 0, 20, desc.background, 1, flags=flags))
        assert not lin.any() and st.rays == 0 and st.paths == 0
        lin, _, st = sc.render(ctx, cam, trt.make_params(16, 8, 1, 1, desc.background, 1, flags=flags))      # one sample, one bounce
        assert st.rays == st.paths == 128


def test_reference_api_renderer(trt, orc, ctx):
    """The reference-shaped API end to end: /root/reference/src/main.rs translated (smaller image)."""
    world = trt.World.new()
    world.add_material("red", trt.Lambertian(trt.Vec3.new(0.65, 0.05, 0.05)))
    world.add_material("white", trt.Lambertian(trt.Vec3.new(0.73, 0.73, 0.73)))
    world.add_material("green", trt.Lambertian(trt.Vec3.new(0.12, 0.45, 0.15)))
    world.add_material("light", trt.Light(trt.Vec3.new_diagonal(15.0)))
    q = lambda c, u, v, m: world.add_geometry(trt.Quad(trt.Vec3.new(*c), trt.Vec3.new(*u), trt.Vec3.new(*v), world.get_material(m)))
    q((100, 0, 0), (0, 100, 0), (0, 0, 100), "green"); q((0, 0, 0), (0, 100, 0), (0, 0, 100), "red")
    q((65, 100, 60), (-30, 0, 0), (0, 0, -20), "light"); q((0, 0, 0), (100, 0, 0), (0, 0, 100), "white")
    q((100, 100, 100), (-100, 0, 0), (0, 0, -100), "white"); q((0, 0, 100), (100, 0, 0), (0, 100, 0), "white")
    world.add_geometry(trt.new_box(trt.Vec3.new(25, 0, 50), trt.Vec3.new(55, 60, 80), world.get_material("white")))
    world.add_geometry(trt.new_box(trt.Vec3.new(45, 0, 10), trt.Vec3.new(75, 30, 40), world.get_material("white")))
    camera = trt.Camera.new(140.0, 0.6, trt.Vec3.new(50, 50, -140), trt.Vec3.new(50, 50, 0), trt.Vec3.new(0, 1, 0), 40.0, 120, 120)
    image = trt.Renderer.new(8, 8, 20, True, trt.Vec3.new_diagonal(0.001), seed=1).render(camera, world)
    desc = trt.scenes.cornell(120, 120, 8, 20)
    osc, ocam, _ = both(trt, orc, desc)
    olin, ou8, _, _ = osc.render(ocam, orc.make_params(120, 120, 8, 20, desc.background, 1))
    assert ulp_diff(image.linear, olin).max() <= 1 and np.abs(image.rgb8.astype(int) - ou8.astype(int)).max() <= 1


def test_cpp_example_renders_main_rs_on_the_gpu(trt, orc, ctx, tmp_path):
    """examples/cornell.cpp is /root/reference/src/main.rs against include/trt.hpp (300x300, 300 spp, depth 20). Compiled,
    run on the GPU, its PNG decoded: the 8-bit image must equal the oracle's render of the same scene within +-1 LSB."""
    import struct, subprocess, zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "tiny-raytracer_b200")
    exe, png = str(tmp_path / "cornell"), str(tmp_path / "out.png")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "cornell.cpp"),
                           "-o", exe, "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    r = subprocess.run([exe, png], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "27000000 paths" in r.stdout, r.stdout + r.stderr
    b = open(png, "rb").read()
    pos, idat = 8, b""
    while pos < len(b):
        n, = struct.unpack(">I", b[pos:pos + 4])
        if b[pos + 4:pos + 8] == b"IDAT": idat += b[pos + 8:pos + 8 + n]
        pos += 12 + n
    img = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(300, 1 + 3 * 300)[:, 1:].reshape(300, 300, 3)
    desc = trt.scenes.cornell(300, 300, 300, 20)
    osc, ocam, _ = both(trt, orc, desc)
    olin, ou8, _, _ = osc.render(ocam, orc.make_params(300, 300, 300, 20, desc.background, 1))
    assert np.abs(img.astype(int) - ou8.astype(int)).max() <= 1 and (img != ou8).mean() < 0.02


# ------------------------------------------------------------------ progressive / resume, streaming Sampler adapter
def test_progressive_render_and_checkpoint_resume(trt, orc, ctx):
    """Samples added in three instalments, with a checkpoint carried into a fresh renderer in between, give the image of
    one pass over the same sample indices (<= 1 ulp) and match the oracle."""
    from tiny_raytracer_b200.progressive import ProgressiveRenderer
    desc = trt.scenes.cornell_spheres(120, 80, 12, 20)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, trt.BVH_AUTO)
    pr = ProgressiveRenderer(ctx, sc, cam, desc.max_bounces, desc.background, desc.seed)
    pr.add_samples(3)
    preview = pr.image()
    o3, _, _, _ = osc.render(ocam, orc.make_params(120, 80, 3, 20, desc.background, desc.seed))
    assert ulp_diff(preview.linear, o3).max() <= 1                      # a 3-spp preview is itself a correct 3-spp render
    pr.add_samples(4)
    ck = pr.state()
    pr2 = ProgressiveRenderer(ctx, sc, cam, desc.max_bounces, desc.background, desc.seed)
    pr2.load_state(ck)
    pr2.add_samples(5)
    assert pr2.samples_done == 12
    img = pr2.image()
    o12, ou8, _, _ = osc.render(ocam, orc.make_params(120, 80, 12, 20, desc.background, desc.seed))
    assert ulp_diff(img.linear, o12).max() <= 1 and np.abs(img.rgb8.astype(int) - ou8.astype(int)).max() <= 1
    with pytest.raises(ValueError):
        ProgressiveRenderer(ctx, sc, cam, desc.max_bounces, desc.background, desc.seed + 1).load_state(ck)


def test_streaming_sampler_adapter(trt, orc, ctx):
    """trt_sampler_*: 25 000 SamplePoints streamed in batches of 4096 with two batches in flight come back complete, in
    order, and equal to one trt_sample_batch call and to the oracle (the reference's liveness test, scaled up)."""
    desc = trt.scenes.five_spheres(64, 48, 1, 10)
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, trt.BVH_AUTO)
    n = 25_000
    rng = np.random.default_rng(5)
    pts = np.zeros(n, trt.SAMPLE_POINT_DTYPE)
    pts["x"] = rng.integers(0, 64, n); pts["y"] = rng.integers(0, 48, n)
    for i in range(n):
        pts["origin"][i] = orc.primary_ray(ocam, int(pts["x"][i]), int(pts["y"][i]), desc.seed, i)[:3]
    d = rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.05) + np.float32([2.0, -2.0, -2.0]) / np.float32(3.4641)
    pts["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    p = trt.make_params(64, 48, 1, 10, desc.background, desc.seed)
    whole = sc.sample_batch(ctx, pts, p)
    st = trt.SamplerStream(ctx, sc, p, batch_capacity=4096)
    chunks = [pts[i:i + 3000] for i in range(0, n, 3000)]                # caller batches need not match the capacity
    got = np.concatenate(list(st.sampling(chunks)))
    st.close()
    assert got.shape[0] == n and np.array_equal(got["x"], pts["x"]) and np.array_equal(got["y"], pts["y"])
    assert np.array_equal(bits(got["color"]), bits(whole["color"]))
    want = osc.sample_batch(pts, orc.make_params(64, 48, 1, 10, desc.background, desc.seed))
    assert np.array_equal(bits(got["color"]), bits(want["color"]))
    # back-pressure is an error value, not a hang
    st = trt.SamplerStream(ctx, sc, p, batch_capacity=128)
    st.submit(pts[:128]); st.submit(pts[128:256])
    with pytest.raises(trt.TrtError):
        st.submit(pts[256:300])
    assert st.collect().shape[0] == 128 and st.collect().shape[0] == 128 and st.collect().shape[0] == 0
    with pytest.raises(trt.TrtError):
        st.submit(pts[:200])                                             # larger than the capacity
    st.close()


def test_device_built_bvh(trt, orc, ctx):
    """TRT_BVH_LBVH_DEVICE: the tree is built on the GPU (Morton + radix sort + Karras) and walked like
    the SAH tree; the reference's own Node::new also runs on the GPU, level by level, and supplies the reference leaf ranks
    (exact-t ties: random_mixed and Cornell are full of them) and the reference tree for the replay of the order-dependent
    rays of DESIGN.md §4.3. So closest-hit ids and t must equal the oracle's bit for bit, like every other kind."""
    for desc, n_rays in ((trt.scenes.random_spheres(200_000, 64, 64, 1, 3), 300_000), (trt.scenes.random_mixed(3000, 17), 200_000),
                         (trt.scenes.final_scene(), 200_000), (trt.scenes.cornell(), 100_000), (trt.scenes.random_mixed(2, 3), 20_000),
                         (trt.scenes.random_mixed(3, 4), 20_000), (trt.scenes.random_mixed(37, 6), 50_000)):
        osc = orc.Scene.from_desc(desc)
        sc = trt.Scene.from_desc(desc, trt.BVH_LBVH_DEVICE)
        sc.upload(ctx)
        info, times = sc.bvh_info(), sc.build_times()
        assert info["nodes"] == desc.n_prims - 1 and 0 < info["max_depth"] <= 64 and times["device_build_ms"] > 0
        rng = np.random.default_rng(7)
        lo = np.maximum(desc.prim_data[:, :3].min(0) - 3, -120); hi = np.minimum(desc.prim_data[:, :3].max(0) + 3, 120)
        o = rng.uniform(lo, hi, (n_rays, 3)).astype(np.float32)
        d = rng.normal(size=(n_rays, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        prim, t = sc.trace_rays(ctx, o, d, 0.001)
        oprim, ot = osc.trace_rays(o, d, 0.001, use_bvh=True)
        assert np.array_equal(bits(t), bits(ot)), "%s: %d t mismatches" % (desc.name, int((bits(t) != bits(ot)).sum()))
        assert np.array_equal(prim, oprim), "%s: %d id mismatches" % (desc.name, int((prim != oprim).sum()))
        assert (prim != 0xFFFFFFFF).mean() > 0.02
    # whole renders through the megakernel on the device-built tree: tie-heavy Cornell, and deep paths through 50 k spheres
    for desc in (trt.scenes.cornell(160, 160, 8, 20), trt.scenes.random_spheres(50_000, 160, 90, 4, 12)):
        osc, ocam, cam = both(trt, orc, desc)
        p, op = params(trt, orc, desc)
        olin, ou8, ost, _ = osc.render(ocam, op)
        lin, u8, st = trt.Scene.from_desc(desc, trt.BVH_LBVH_DEVICE).render(ctx, cam, p)
        assert st.paths == ost.paths and st.rays == ost.rays
        assert ulp_diff(lin, olin).max() <= 1 and psnr8(u8, ou8) >= PSNR_MIN_DB


@pytest.mark.parametrize("scene", ["mixed2", "mixed3", "mixed5", "mixed800", "cornell", "final", "spheres30k", "dups"])
def test_device_node_new_gives_the_reference_leaf_order(trt, orc, ctx, scene):
    """Node::new (bvh.rs:42-84) run level by level on the GPU must produce exactly the leaf order of the host builder and of
    the oracle's independent builder: same longest-axis choice, stable sort with total_cmp, unsorted pairs, split at n/2.
    `dups` is 600 primitives drawn from 40 distinct ones, so the sort is full of equal keys and only stability decides."""
    S = trt.scenes
    desc = {"mixed2": lambda: S.random_mixed(2, 3), "mixed3": lambda: S.random_mixed(3, 4), "mixed5": lambda: S.random_mixed(5, 8),
            "mixed800": lambda: S.random_mixed(800, 4), "cornell": S.cornell, "final": S.final_scene,
            "spheres30k": lambda: S.random_spheres(30_000, 64, 64, 1, 3), "dups": lambda: S.random_mixed(40, 12)}[scene]()
    if scene == "dups":
        idx = np.random.default_rng(5).integers(0, 40, 600)
        desc = S.SceneDesc(desc.name, desc.materials, desc.prim_type[idx], desc.prim_data[idx], desc.prim_mat[idx], dict(desc.camera),
                           desc.spp, desc.max_bounces, desc.background, desc.seed)
    host = trt.Scene.from_desc(desc, trt.BVH_REFERENCE).leaf_ranks()
    assert np.array_equal(host, orc.Scene.from_desc(desc).leaf_ranks())
    sc = trt.Scene.from_desc(desc, trt.BVH_LBVH_DEVICE)
    sc.upload(ctx)
    dev = sc.leaf_ranks()
    assert np.array_equal(np.sort(dev), np.arange(desc.n_prims)) and np.array_equal(dev, host), "%d ranks differ" % int((dev != host).sum())


def test_big_sah_scene_takes_its_reference_order_from_the_device(trt, orc, ctx):
    """From 65 536 primitives on, TRT_BVH_SAH builds only the SAH tree on the host; ranks and the replay tree come from the
    device Node::new during upload. Same ranks as the host reference builder, and rays equal to the oracle."""
    desc = trt.scenes.random_spheres(70_000, 64, 64, 1, 3)
    host = trt.Scene.from_desc(desc, trt.BVH_REFERENCE).leaf_ranks()
    sc = trt.Scene.from_desc(desc, trt.BVH_SAH)
    assert np.array_equal(sc.leaf_ranks(), np.arange(desc.n_prims))       # placeholder before the upload
    sc.upload(ctx)
    assert np.array_equal(sc.leaf_ranks(), host) and sc.build_times()["device_build_ms"] > 0
    rng = np.random.default_rng(11)
    o = rng.uniform(-60, 60, (200_000, 3)).astype(np.float32)
    d = rng.normal(size=(200_000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    prim, t = sc.trace_rays(ctx, o, d, 0.001)
    oprim, ot = orc.Scene.from_desc(desc).trace_rays(o, d, 0.001, use_bvh=True)
    assert np.array_equal(prim, oprim) and np.array_equal(bits(t), bits(ot))


@pytest.mark.parametrize("n", [2, 31, 32, 33, 63, 64])
def test_flat_list_mask_boundaries(trt, orc, ctx, n):
    """TRT_BVH_FLAT keeps its candidates in two 32-bit masks: primitive counts around 32 and at the limit of 64, on a
    scene full of coplanar and duplicate primitives (exact-t ties), whole paths bit-exact; 65 primitives are refused."""
    desc = trt.scenes.random_mixed(n, 1, 96, 64, 3, 10)
    assert desc.n_prims == n
    osc, ocam, cam = both(trt, orc, desc)
    sc = trt.Scene.from_desc(desc, trt.BVH_FLAT)
    p, op = params(trt, orc, desc)
    prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
    oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True)
    assert np.array_equal(prim, oprim) and np.array_equal(bits(t), bits(ot)) and np.array_equal(nr, onr) and np.array_equal(bits(rad), bits(orad))
    lin, u8, st = sc.render(ctx, cam, p)
    olin, ou8, ost, _ = osc.render(ocam, op)
    assert st.rays == ost.rays and ulp_diff(lin, olin).max() <= 1
    with pytest.raises(trt.TrtError):
        trt.Scene.from_desc(trt.scenes.random_mixed(65, 1), trt.BVH_FLAT)
    assert trt.Scene.from_desc(trt.scenes.random_mixed(65, 1), trt.BVH_AUTO).bvh_info()["nodes"] == 64


@pytest.mark.parametrize("log2_scale", [-20, 0, 20, 28, 40])
@pytest.mark.parametrize("scene", ["cornell_spheres", "mixed40"])
def test_flat_list_conservative_box_phase_extremes(trt, orc, ctx, scene, log2_scale):
    """Phase A of TRT_BVH_FLAT is a conservative box test (centre/half-width, FMA) that is only taken while nothing can
    overflow: |coordinate| <= 2^30, |origin| < 2^30, |1/d| < 2^64; everything else replays the reference scan. Scenes
    scaled by powers of two (exact) across those limits, with rays whose 1/d is huge, must stay bit-exact."""
    S = trt.scenes
    base = S.cornell_spheres() if scene == "cornell_spheres" else S.random_mixed(40, 9)
    f = np.float32(2.0 ** log2_scale)
    desc = S.SceneDesc(base.name, base.materials, base.prim_type, (base.prim_data * f).astype(np.float32), base.prim_mat,
                       dict(base.camera), base.spp, base.max_bounces, base.background, base.seed)
    osc = orc.Scene.from_desc(desc)
    rng = np.random.default_rng(107 + log2_scale)
    n = 60_000
    lo = base.prim_data[:, :3].min(0) - 3; hi = base.prim_data[:, :3].max(0) + 3
    o = (rng.uniform(lo, hi, (n, 3)).astype(np.float32) * f).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    d[:3000, 0] = np.float32(1e-22); d[3000:6000, 1] = np.float32(-3e-30); d[6000:7000, 2] = np.float32(1e-40)   # |1/d| up to inf
    d[7000:8000, 0] = 0.0
    o[8000:9000] = np.float32(3e9)                                                                            # |o| > 2^30
    o[9000:12000] = np.round(o[9000:12000] / f) * f                                                           # origins on planes
    t_min = np.float32(0.001) * f
    oprim, ot = osc.trace_rays(o, d, float(t_min), use_bvh=True)
    for kind in (trt.BVH_FLAT, trt.BVH_REFERENCE):
        prim, t = trt.Scene.from_desc(desc, kind).trace_rays(ctx, o, d, float(t_min))
        assert np.array_equal(prim, oprim), "kind %d: %d prim mismatches" % (kind, int((prim != oprim).sum()))
        assert np.array_equal(bits(t), bits(ot))
    assert (oprim != 0xFFFFFFFF).sum() > n // 50


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_cuda_matches_committed_golden_fixtures(trt, ctx, kind):
    """The CUDA path against tests/golden/golden_small.npz (oracle outputs committed with their generator): no oracle run
    is involved here, so this also holds if the oracle library cannot be built on the GPU box."""
    import importlib.util, os
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gdir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    g = np.load(os.path.join(gdir, "golden_small.npz"))
    for name, mk in mg.CASES.items():
        d = mk(trt.scenes); c = d.camera
        sc = trt.Scene.from_desc(d, kind)
        cam = trt.make_camera(**c)
        p = trt.make_params(c["width"], c["height"], d.spp, d.max_bounces, d.background, d.seed)
        prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
        lin, u8, st = sc.render(ctx, cam, p)
        assert np.array_equal(prim, g[name + "/prim"]) and np.array_equal(bits(t), bits(g[name + "/t"])), name
        assert np.array_equal(bits(rad), bits(g[name + "/rad"])) and np.array_equal(nr, g[name + "/nr"]), name
        assert ulp_diff(lin, g[name + "/lin"]).max() <= 1 and np.abs(u8.astype(int) - g[name + "/u8"].astype(int)).max() <= 1, name
        assert (st.paths, st.rays) == (int(g[name + "/counts"][0]), int(g[name + "/counts"][1])), name
        assert np.array_equal(sc.leaf_ranks(), g[name + "/ranks"]), name


def test_differential_fuzz_campaign(trt, ctx):
    """scripts/fuzz_parity.py: random mixed worlds of 1..3000 primitives, random cameras (with and without a lens), depths
    and seeds, every BVH kind, binary / wide / step-wise walks: primary hits, path lengths, per-sample radiance, ray counts
    and images equal to the oracle bit for bit, plus arbitrary and degenerate rays through trt_trace_rays. 120 scenes here;
    9 500 scenes ran clean during the round."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_parity.py"), "120", "777"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]



# ------------------------------------------------------------------ round 2: the gaps VERDICT r1 named
def _one_prim_desc(trt, prim_kind, data):
    d = np.zeros((1, 12), np.float32); d[0, :len(data)] = data
    return trt.scenes.SceneDesc("one", [(0, (0.5, 0.5, 0.5), 0.0)], np.array([prim_kind], np.uint32), d, np.zeros(1, np.uint32))


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
def test_reference_kat_rays_through_the_cuda_path(trt, orc, ctx, kind):
    """The reference's OWN known-answer rays — hittable/sphere.rs:68-100 (test_sphere_hit) and hittable/quad.rs:67-96
    (test_quad_hit) — pushed through trt_trace_rays for every BVH kind (round 1 ran them through the oracle only):
    same hit / miss, t equal to the reference's expected value under the reference's own tolerance, and equal to the
    oracle bit for bit."""
    import math
    norm = lambda v: (np.float32(v) / np.sqrt(np.float32(np.dot(np.float32(v), np.float32(v))))).astype(np.float32)   # Ray::new, ray.rs:12-14
    # sphere (0,0,-1) r 0.5: straight hit t = 0.5; oblique hit t ~ sqrt(3)/2 (1e-2, as the reference asserts); miss
    sc = trt.Scene.from_desc(_one_prim_desc(trt, 0, [0, 0, -1, 0.5]), kind)
    osc = orc.Scene.from_desc(_one_prim_desc(trt, 0, [0, 0, -1, 0.5]))
    o = np.zeros((3, 3), np.float32)
    d = np.stack([norm([0, 0, -1]), norm([0, math.sqrt(3.0), -3.0]), norm([0, -1, -1])])
    prim, t = sc.trace_rays(ctx, o, d, 0.0)
    oprim, ot = osc.trace_rays(o, d, 0.0, use_bvh=True)
    assert list(prim) == [0, 0, 0xFFFFFFFF] and t[0] == 0.5 and abs(t[1] - math.sqrt(3) / 2) < 1e-2
    assert np.array_equal(prim, oprim) and np.array_equal(bits(t), bits(ot))
    p1 = o[1] + t[1] * d[1]
    assert np.linalg.norm(p1 - np.float32([0, math.sqrt(3) / 4, -0.75])) < 1e-2
    # quad corner 0, u = x, v = 3z: hit from above t = 1; diagonal hit t = sqrt 2 (1e-6); a ray that leaves the quad's extent
    sc = trt.Scene.from_desc(_one_prim_desc(trt, 1, [0, 0, 0, 1, 0, 0, 0, 0, 3]), kind)
    osc = orc.Scene.from_desc(_one_prim_desc(trt, 1, [0, 0, 0, 1, 0, 0, 0, 0, 3]))
    o = np.float32([[0, 1, 0], [0, 1, 1], [0, 1, 1]])
    d = np.stack([norm([0, -1, 0]), norm([0, -1, -1]), norm([0, -1, -1.1])])
    prim, t = sc.trace_rays(ctx, o, d, 0.0)
    oprim, ot = osc.trace_rays(o, d, 0.0, use_bvh=True)
    assert list(prim) == [0, 0, 0xFFFFFFFF] and t[0] == 1.0 and abs(t[1] - math.sqrt(2)) < 1e-6
    assert np.all(np.abs(o[1] + t[1] * d[1]) < 1e-6)
    assert np.array_equal(prim, oprim) and np.array_equal(bits(t), bits(ot))


@pytest.mark.parametrize("kind", [3, 1])     # TRT_BVH_AUTO (= trees built on the GPU at this size) and the host-built SAH tree
def test_c4_million_spheres_matches_oracle(trt, orc, ctx, kind):
    """BASELINE config 4 at its REAL scene size — 10^6 random spheres, depth 50 — at a quarter of the resolution and one
    sample per pixel so that the oracle finishes in seconds: primary hits, path lengths and per-sample radiance bit-exact,
    ray counts equal, image <= 1 ulp (whole-pixel items: in fact bit-exact)."""
    desc = trt.scenes.random_spheres(1_000_000, 480, 270, 1, 50)
    osc, ocam, cam = both(trt, orc, desc)
    p, op = params(trt, orc, desc, flags=2)
    oprim, ot, orad, onr = osc.render_aov(ocam, op, True, True)
    olin, ou8, ost, _ = osc.render(ocam, op)
    sc = trt.Scene.from_desc(desc, kind)
    sc.upload(ctx)
    if kind == 3:
        assert sc.build_times()["device_build_ms"] > 0          # AUTO resolved to the device build
    prim, t, rad, nr = sc.render_aov(ctx, cam, p, True, True)
    same = (nr == onr) & (bits(rad) == bits(orad)).all(1) & (prim == oprim) & (bits(t) == bits(ot))
    assert same.all(), "%d of %d samples differ" % (int((~same).sum()), same.size)
    lin, u8, st = sc.render(ctx, cam, p)
    assert (st.paths, st.rays) == (ost.paths, ost.rays)
    assert ulp_diff(lin, olin).max() <= 1 and psnr8(u8, ou8) >= PSNR_MIN_DB
    assert st.rays > 3 * st.paths                               # deep paths really happen in this scene


def test_wavefront_and_event_counting_replay_order_dependent_rays(trt, orc, ctx):
    """the wavefront extend kernel and the event-counting build of the walk skipped the replay of
    order-dependent rays, so TRT_VARIANT_WAVEFRONT / TRT_FLAG_COUNT_EVENTS could return another hit than
    the megakernel for a few rays per million. Same scene as test_sah_order_dependent_rays_are_replayed (200 k spheres,
    depth 50, 1 M samples, where 3-5 such rays occur): every variant must give the oracle's ray count and image."""
    desc = trt.scenes.random_spheres(200_000, 960, 540, 2, 50)
    osc, ocam, cam = both(trt, orc, desc)
    _, op = params(trt, orc, desc)
    olin, ou8, ost, _ = osc.render(ocam, op)
    w, h = desc.camera["width"], desc.camera["height"]
    for kind in (trt.BVH_SAH, trt.BVH_LBVH_DEVICE):
        sc = trt.Scene.from_desc(desc, kind)
        for variant, flags in ((trt.VARIANT_MEGAKERNEL, 2), (trt.VARIANT_MEGAKERNEL, 2 | 1), (trt.VARIANT_WAVEFRONT, 0), (trt.VARIANT_WAVEFRONT, 4)):
            p = trt.make_params(w, h, desc.spp, desc.max_bounces, desc.background, desc.seed, variant=variant, flags=flags)
            lin, u8, st = sc.render(ctx, cam, p)
            assert (st.paths, st.rays) == (ost.paths, ost.rays), (kind, variant, flags, st.rays, ost.rays)
            d = ulp_diff(lin, olin)
            assert d.max() <= 1, "kind %d variant %d flags %d: %d values differ by up to %d ulp" % (kind, variant, flags, int((d > 1).sum()), int(d.max()))


def test_scene_recommit_while_streaming_sampler_is_alive(trt, orc, ctx):
    """a scene re-committed (and re-uploaded) while a streaming sampler holds its device copy. The next submit
    must pick up the new copy — never read freed or half-written memory — and results stay equal to the oracle."""
    desc = trt.scenes.five_spheres(64, 48, 1, 10)
    osc = orc.Scene.from_desc(desc)
    sc = trt.Scene.from_desc(desc, trt.BVH_SAH)
    rng = np.random.default_rng(3)
    pts = np.zeros(6000, trt.SAMPLE_POINT_DTYPE)
    pts["x"] = rng.integers(0, 64, 6000); pts["y"] = rng.integers(0, 48, 6000)
    pts["origin"] = np.float32([-2.0, 2.0, 1.0])
    d = rng.normal(size=(6000, 3)).astype(np.float32) * np.float32(0.08) + np.float32([2.0, -2.0, -2.0]) / np.float32(3.4641)
    pts["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    p = trt.make_params(64, 48, 1, 10, desc.background, desc.seed)
    want = osc.sample_batch(pts, orc.make_params(64, 48, 1, 10, desc.background, desc.seed))
    st = trt.SamplerStream(ctx, sc, p, batch_capacity=2000)
    st.submit(pts[:2000]); st.submit(pts[2000:4000])
    sc.commit(trt.BVH_REFERENCE); sc.upload(ctx)                # re-commit with another kind while two batches are in flight
    a = st.collect(); b = st.collect()
    st.submit(pts[4000:6000])                                   # served from the new device copy
    c = st.collect()
    st.close()
    got = np.concatenate([a, b, c])
    assert np.array_equal(bits(got["color"]), bits(want["color"]))


def test_two_gpus_equal_one_gpu_equal_oracle(trt, orc, ctx):
    """N-GPU parity on hardware (skipped on a one-GPU box): two trt_ctx on two devices, each renders its share of the
    samples (multigpu.partition_samples), the f64 sums are added, and the image must equal the 1-GPU image and the
    oracle (<= 1 ulp, ray counts exactly) — through the Python plumbing AND through trt_render_multi (C ABI)."""
    import torch
    from tiny_raytracer_b200.multigpu import partition_samples
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    for desc in (trt.scenes.cornell(320, 200, 24, 20), trt.scenes.final_scene(160, 90, 7, 50)):
        osc, ocam, cam = both(trt, orc, desc)
        w, h = desc.camera["width"], desc.camera["height"]
        p, op = params(trt, orc, desc)
        olin, ou8, ost, _ = osc.render(ocam, op)
        sc = trt.Scene.from_desc(desc, trt.BVH_AUTO)
        one, one8, st1 = sc.render(ctx, cam, p)
        ctxs = [ctx, trt.Context(1)]
        accs, rays = [], 0
        for r, c in enumerate(ctxs):
            off, cnt = partition_samples(desc.spp, 2, r)
            acc = torch.zeros(h * w * 3, dtype=torch.float64, device="cuda:%d" % r)
            torch.cuda.synchronize(r)
            pr, _ = params(trt, orc, desc, spp=cnt, sample_offset=off, spp_total=desc.spp)
            s = sc.render_accumulate(c, cam, pr, acc.data_ptr())
            c.synchronize()
            rays += int(s.rays); accs.append(acc)
        total = accs[0] + accs[1].to("cuda:0")
        lin = (total * (1.0 / desc.spp)).to(torch.float32).cpu().numpy().reshape(h, w, 3)
        assert rays == st1.rays == ost.rays
        assert ulp_diff(lin, one).max() <= 1 and ulp_diff(lin, olin).max() <= 1
        # the same through the C ABI: one call, both devices, P2P reduce-scatter + tone map on the GPUs
        mlin, mu8, mst = sc.render_multi([0, 1], cam, p)
        assert (mst.paths, mst.rays) == (ost.paths, ost.rays)
        assert ulp_diff(mlin, one).max() <= 1 and ulp_diff(mlin, olin).max() <= 1
        assert np.abs(mu8.astype(int) - ou8.astype(int)).max() <= 1 and psnr8(mu8, ou8) >= PSNR_MIN_DB
        ctxs[1].close()
    # trees built on both devices at once (parallel uploads inside trt_multi_render)
    big = trt.scenes.random_spheres(70_000, 96, 54, 4, 8)
    osc, ocam, cam = both(trt, orc, big)
    p, op = params(trt, orc, big)
    olin, _, ost, _ = osc.render(ocam, op)
    lin, u8, st = trt.Scene.from_desc(big, trt.BVH_AUTO).render_multi([0, 1], cam, p)
    assert (st.paths, st.rays) == (ost.paths, ost.rays) and ulp_diff(lin, olin).max() <= 1


def test_render_multi_on_one_device_equals_render(trt, orc, ctx):
    """trt_render_multi with a device list of length 1 (and with the same device twice: two ctx, two shares, one GPU) goes
    through the same partition + reduce-scatter + tone-map path as the 8-GPU run and must equal trt_render and the oracle."""
    desc = trt.scenes.cornell_spheres(200, 120, 10, 20)
    osc, ocam, cam = both(trt, orc, desc)
    p, op = params(trt, orc, desc)
    olin, ou8, ost, _ = osc.render(ocam, op)
    sc = trt.Scene.from_desc(desc, trt.BVH_AUTO)
    one, one8, st1 = sc.render(ctx, cam, p)
    for devs in ([0], [0, 0], [0, 0, 0]):
        lin, u8, st = sc.render_multi(devs, cam, p)
        assert (st.paths, st.rays) == (ost.paths, ost.rays) == (st1.paths, st1.rays)
        assert ulp_diff(lin, olin).max() <= 1 and ulp_diff(lin, one).max() <= 1
        assert np.abs(u8.astype(int) - ou8.astype(int)).max() <= 1
    with pytest.raises(trt.TrtError):
        sc.render_multi([], cam, p)
    with pytest.raises(trt.TrtError):
        sc.render_multi([99], cam, p)
    # fewer samples than shares: one share is empty and still takes part in the reduce
    p2, op2 = params(trt, orc, desc, spp=2)
    o2, _, ost2, _ = osc.render(ocam, op2)
    lin, u8, st = sc.render_multi([0, 0, 0], cam, p2)
    assert (st.paths, st.rays) == (ost2.paths, ost2.rays) and ulp_diff(lin, o2).max() <= 1
    # a persistent handle renders several frames and several scenes (buffers and contexts are reused); trees built on the device
    mc = trt.MultiContext([0, 0])
    big = trt.scenes.random_spheres(70_000, 96, 54, 3, 8)
    bosc, bocam, bcam = both(trt, orc, big)
    bp, bop = params(trt, orc, big)
    bolin, _, bost, _ = bosc.render(bocam, bop)
    bsc = trt.Scene.from_desc(big, trt.BVH_AUTO)                 # >= 65 536 primitives: LBVH + reference order built during the upload
    for _ in range(2):
        lin, u8, st = mc.render(bsc, bcam, bp)
        assert (st.paths, st.rays) == (bost.paths, bost.rays) and ulp_diff(lin, bolin).max() <= 1
        lin, u8, st = mc.render(sc, cam, p)
        assert st.rays == ost.rays and ulp_diff(lin, olin).max() <= 1
    assert np.array_equal(bsc.leaf_ranks(), orc.Scene.from_desc(big).leaf_ranks())
    mc.close()


def test_cpp_multi_gpu_example_equals_one_gpu(trt, orc, ctx, tmp_path):
    """examples/cornell_multi.cpp: a plain C++ host (no Python, no NCCL) renders the Cornell box on ALL GPUs of the box through
    trt::Renderer -> trt_render_multi and on one GPU through trt_render; it exits 0 only if ray counts are equal and the images
    agree to <= 1 ulp / <= 1 LSB. (On a one-GPU box the device list has one entry and both renders take the single-GPU path.)"""
    import json, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "tiny-raytracer_b200")
    exe = str(tmp_path / "cornell_multi")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "cornell_multi.cpp"),
                           "-o", exe, "-L" + pkg, "-ltrt_b200", "-Wl,-rpath," + pkg])
    r = subprocess.run([exe, "0", "480", "270", "64"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["rays_1gpu"] == d["rays_ngpu"] and d["linear_max_ulp_vs_1gpu"] <= 1 and d["rgb8_max_lsb_vs_1gpu"] <= 1
    desc = trt.scenes.cornell(480, 270, 64, 50)
    _, _, ost, _ = orc.Scene.from_desc(desc).render(orc.make_camera(**desc.camera), orc.make_params(480, 270, 64, 50, desc.background, 1))
    assert d["rays_ngpu"] == ost.rays


def test_instanced_geometry_renders_like_explicit_geometry(trt, orc, ctx):
    """`Transform`-instanced geometry. The Cornell box with its two boxes placed as INSTANCES of one
    template box (translation only, exact in f32) is the same scene as main.rs builds explicitly: bit-identical to the oracle.
    A rotated + uniformly scaled instance of a sphere-and-box template renders like the same primitives added explicitly."""
    T = trt.Transform
    def walls(world):
        world.add_material("red", trt.Lambertian((0.65, 0.05, 0.05))); world.add_material("white", trt.Lambertian((0.73, 0.73, 0.73)))
        world.add_material("green", trt.Lambertian((0.12, 0.45, 0.15))); world.add_material("light", trt.Light((15.0, 15.0, 15.0)))
        q = lambda c, u, v, m: world.add_geometry(trt.Quad(c, u, v, world.get_material(m)))
        q((100, 0, 0), (0, 100, 0), (0, 0, 100), "green"); q((0, 0, 0), (0, 100, 0), (0, 0, 100), "red")
        q((65, 100, 60), (-30, 0, 0), (0, 0, -20), "light"); q((0, 0, 0), (100, 0, 0), (0, 0, 100), "white")
        q((100, 100, 100), (-100, 0, 0), (0, 0, -100), "white"); q((0, 0, 100), (100, 0, 0), (0, 100, 0), "white")
    world = trt.World.new(); walls(world)
    white = world.get_material("white")
    world.add_instance(trt.new_box((0, 0, 0), (30, 60, 30), white), T.translate((25, 0, 50)))
    world.add_instance(trt.new_box((0, 0, 0), (30, 30, 30), white), T.translate((45, 0, 10)))
    camera = trt.Camera.new(140.0, 0.6, (50, 50, -140), (50, 50, 0), (0, 1, 0), 40.0, 96, 96)
    img = trt.Renderer.new(8, 8, 20, True, (0.001, 0.001, 0.001), seed=1).render(camera, world)
    desc = trt.scenes.cornell(96, 96, 8, 20)
    olin, ou8, _, _ = orc.Scene.from_desc(desc).render(orc.make_camera(**desc.camera), orc.make_params(96, 96, 8, 20, desc.background, 1))
    assert np.array_equal(bits(img.linear), bits(olin))
    # a general similarity: the instance equals the explicitly transformed primitives (same host arithmetic), and the oracle agrees
    t = T.new((50, 20, 50), (0.4, 0.7, -0.2), (1.5, 1.5, 1.5))
    glass = trt.Dielectric((1.0, 1.0, 1.0), 1.5)
    template = trt.new_box((-5, -5, -5), (5, 5, 5), white) + [trt.Sphere((0, 9, 0), 4.0, glass)]
    w2 = trt.World.new(); walls(w2); w2.add_material("glass", glass); w2.add_instance(template, t)
    d2 = w2.to_desc()
    assert d2.n_prims == 13 and abs(float(d2.prim_data[12, 3]) - 6.0) < 1e-5
    d2 = trt.scenes.SceneDesc("instanced", d2.materials, d2.prim_type, d2.prim_data, d2.prim_mat, dict(desc.camera), 6, 20, desc.background, 3)
    osc, ocam, cam = both(trt, orc, d2)
    p, op = params(trt, orc, d2)
    lin, u8, st = trt.Scene.from_desc(d2, trt.BVH_AUTO).render(ctx, cam, p)
    olin, ou8, ost, _ = osc.render(ocam, op)
    assert st.rays == ost.rays and ulp_diff(lin, olin).max() <= 1


@pytest.mark.parametrize("kind", [0, 1, 2, 4])
def test_first_bounce_pass_is_bit_identical_to_inline_primary_rays(trt, orc, ctx, kind):
    """Round 2: primary rays can be generated and traced by a coherent pass (k_primary) that hands the path loop one record per
    path — the default for wide-node trees in shared memory (where it was measured to pay), forced by TRT_FLAG_PRIMARY_PASS (32)
    and forbidden by TRT_FLAG_INLINE_PRIMARY (16). Same arithmetic, same Philox streams, same summation order: images
    (whole-pixel items) must be bit-identical to each other and to the oracle, ray / path counts equal — also when the
    samples are rendered in several batches because the records do not fit the scratch budget, with T-then-S and step-wise
    loops, with a lens and without, when every path ends at its first segment (max_bounces 1)."""
    S = trt.scenes
    cases = [S.cornell_spheres(96, 64, 12, 20), S.final_scene(80, 45, 9, 50), S.glass_cornell(64, 36, 10, 100), S.random_mixed(300, 7, 72, 40, 6, 12),
             S.cornell(50, 30, 7, 1), S.five_spheres(64, 48, 5, 10)]
    for desc in cases:
        if kind == 2 and desc.n_prims > 64:
            continue
        osc, ocam, cam = both(trt, orc, desc)
        sc = trt.Scene.from_desc(desc, kind)
        _, op = params(trt, orc, desc)
        olin, ou8, ost, _ = osc.render(ocam, op)
        for regen in (0, 9):
            out = {}
            for name, flags, budget in (("auto", 2, None), ("pass", 2 | 32, None), ("inline", 2 | 16, None), ("batched", 2 | 32, "1"), ("tail_split", 32, "1")):
                if budget is None: os.environ.pop("TRT_PRIMARY_BUDGET_MB", None)
                else: os.environ["TRT_PRIMARY_BUDGET_MB"] = budget          # 1 MB of records: a few samples per batch
                p, _ = params(trt, orc, desc, flags=flags | (regen << 8))
                out[name] = sc.render(ctx, cam, p)
            os.environ.pop("TRT_PRIMARY_BUDGET_MB", None)
            for name, (lin, u8, st) in out.items():
                assert (st.paths, st.rays) == (ost.paths, ost.rays), (desc.name, kind, regen, name, st.rays, ost.rays)
                if name == "tail_split": assert ulp_diff(lin, olin).max() <= 1, (desc.name, kind, regen, name)
                else: assert np.array_equal(bits(lin), bits(olin)), (desc.name, kind, regen, name, int((bits(lin) != bits(olin)).sum()))
            w, h = desc.camera["width"], desc.camera["height"]
            launches = {name: out[name][2].launches for name in out}
            assert launches["pass"] == launches["inline"] + 1, (desc.name, kind, regen, launches)      # k_primary + path loop (+ tone map)
            per_batch = (1 << 20) // (w * h * 32)
            if per_batch < desc.spp:          # the 1 MB budget holds fewer samples than the frame has: several batches of two launches
                assert launches["batched"] == 2 * -(-desc.spp // per_batch) + 1, (desc.name, kind, regen, launches)


@pytest.mark.parametrize("kind", [1, 0])
def test_step_functions_of_the_step_wise_loop_agree(trt, orc, ctx, kind):
    """Round 2: the step-wise megakernel loop has three step functions for wide nodes in global memory (trt.h TRT_FLAG_STEP):
    node + leaf + pop in one step (1), one dependent fetch per step (2), and the latter with the top stack entry in registers
    (3). Same tests on the same records in the same per-lane order: images must be bit-identical to each other and to the
    oracle. 70 000 spheres keep the tree out of shared memory (kind 1 = SAH -> ordered walk with replay, kind 0 = the
    reference topology on wide nodes); a mixed scene covers quads; regen 9 forces the step-wise loop on small trees too
    (those stay in shared memory and must ignore the request)."""
    S = trt.scenes
    cases = [S.random_spheres(70000, 96, 54, 3, 50), S.random_mixed(300, 7, 72, 40, 6, 12), S.final_scene(80, 45, 4, 50)]
    for desc in cases:
        osc, ocam, cam = both(trt, orc, desc)
        sc = trt.Scene.from_desc(desc, kind)
        _, op = params(trt, orc, desc)
        olin, ou8, ost, _ = osc.render(ocam, op)
        for step in (1, 2, 3):
            for regen in (0, 9, 20):
                p, _ = params(trt, orc, desc, flags=2 | (regen << 8) | trt._lib.flag_step(step))
                lin, u8, st = sc.render(ctx, cam, p)
                assert (st.paths, st.rays) == (ost.paths, ost.rays), (desc.name, kind, step, regen, st.rays, ost.rays)
                assert np.array_equal(bits(lin), bits(olin)), (desc.name, kind, step, regen, int((bits(lin) != bits(olin)).sum()))


@pytest.mark.parametrize("kind", [1, 3])
def test_compressed_wide8_walk_is_bit_identical_to_the_wide4_walk(trt, orc, ctx, kind):
    """Round 2: ordered trees that do not fit shared memory can be walked on compressed 8-wide nodes (TRT_WIDE8=1 at upload; quantised
    child boxes that only cull; a leaf's own box is rebuilt / fetched and tested exactly). TRT_FLAG_NO_WIDE8 keeps the 4-wide walk.
    Images must be bit-identical to each other and to the oracle, ray counts equal, for the host SAH tree (kind 1; >= 65 536
    primitives: wide records made on the device) and TRT_BVH_AUTO (kind 3; device-built tree from 65 536 primitives), on spheres
    and on a mixed scene with quads, with the step-wise and the T-then-S loop; and the walk must really be the 8-wide one
    (TRT_FLAG_COUNT_WALKED: fewer node records per ray than the 4-wide walk)."""
    S = trt.scenes
    cases = [S.random_spheres(70_000, 96, 54, 3, 50), S.random_mixed(3000, 17, 72, 40, 4, 12), S.random_mixed(70_000, 23, 64, 36, 2, 20)]
    os.environ["TRT_WIDE8"] = "1"          # opt-in, read when the scene is uploaded (measured slower than the 4-wide walk on B200)
    try:
        scenes = []
        for desc in cases:
            sc = trt.Scene.from_desc(desc, kind)
            sc.upload(ctx)
            scenes.append(sc)
    finally:
        os.environ.pop("TRT_WIDE8", None)
    for desc, sc in zip(cases, scenes):
        osc, ocam, cam = both(trt, orc, desc)
        _, op = params(trt, orc, desc)
        olin, ou8, ost, _ = osc.render(ocam, op)
        for regen in (0, 9):
            out = {}
            for name, flags in (("wide8", 2), ("wide4", 2 | trt._lib.FLAG_NO_WIDE8)):
                p, _ = params(trt, orc, desc, flags=flags | (regen << 8))
                out[name] = sc.render(ctx, cam, p)
            for name, (lin, u8, st) in out.items():
                assert (st.paths, st.rays) == (ost.paths, ost.rays), (desc.name, kind, regen, name, st.rays, ost.rays)
                assert np.array_equal(bits(lin), bits(olin)), (desc.name, kind, regen, name, int((bits(lin) != bits(olin)).sum()))
        walked = {}
        for name, flags in (("wide8", 2 | 8), ("wide4", 2 | 8 | trt._lib.FLAG_NO_WIDE8)):
            p, _ = params(trt, orc, desc, flags=flags | (9 << 8))
            lin, u8, st = sc.render(ctx, cam, p)
            assert np.array_equal(bits(lin), bits(olin)), (desc.name, kind, name)
            walked[name] = st.node_visits / max(st.rays, 1)
        assert 0 < walked["wide8"] < 0.9 * walked["wide4"], (desc.name, kind, walked)
        # arbitrary and degenerate rays through trt_trace_rays (one thread per ray, the same wide-8 walk): exact zeros, negative
        # zeros and axis-aligned directions cannot take the quantised-box test and must fall back to the reference-order walk;
        # t_min = 0 and a negative t_min as well
        rng = np.random.default_rng(11)
        n = 60_000
        lo = np.maximum(desc.prim_data[:, :3].min(0) - 3, -120); hi = np.minimum(desc.prim_data[:, :3].max(0) + 3, 120)
        o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        d[:2000, 0] = 0.0; d[2000:4000, 1] = -0.0; d[4000:5000] = np.float32([0, 0, 1]); d[5000:6000] = np.float32([0, -1, 0])
        o[6000:8000] = np.round(o[6000:8000])
        osc_t = orc.Scene.from_desc(desc)
        for t_min in (0.001, 0.0, -1.0):
            prim, t = sc.trace_rays(ctx, o, d, t_min)
            oprim, ot = osc_t.trace_rays(o, d, t_min, use_bvh=True)
            assert np.array_equal(prim, oprim), (desc.name, kind, t_min, int((prim != oprim).sum()))
            assert np.array_equal(bits(t), bits(ot)), (desc.name, kind, t_min)
