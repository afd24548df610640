"""CPU tests: the oracle against every known-answer vector the reference's own tests hold for the hot path
(SURVEY.md §4 / §8c), against Random123 Philox vectors, against brute force, and against the committed golden fixtures."""
import math
import os

import numpy as np
import pytest

F = np.float32
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_small.npz")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---- math/vec3.rs:268-284 vector_test
def test_vec3_kat(orc):
    ln, dt, cr = orc.vec3_kat((1, 2, 3), (4, 5, 6))
    assert F(ln) == np.sqrt(F(14.0)) and dt == 32.0 and list(cr) == [-3.0, 6.0, -3.0]


# ---- ray.rs:33-40 test_at (proves Ray::new normalises)
def test_ray_at_kat(orc):
    r = orc.make_ray((0, 0, 0), (1, 1, -1))
    p = orc.ray_at(r, np.sqrt(F(3.0)))
    assert np.linalg.norm(p - F([1, 1, -1])) < 1e-6
    assert abs(np.linalg.norm(r[3:]) - 1.0) < 1e-6


# ---- camera.rs:78-105 test_new
def test_camera_kat(orc, trt):
    aspect = F(16) / F(9)
    for mk in (orc.make_camera, trt.make_camera):
        cam = mk(1.0, 10.0, (0, 0, 0), (0, 0, 1), (0, 1, 0), 90.0, 16, 9)
        near = lambda a, b: np.all(np.abs(np.array(a[:], np.float32) - F(b)) < 1e-7)     # Vec3's approximate PartialEq, vec3.rs:189-201
        assert near(cam.position, [0, 0, 0]) and near(cam.viewport_upper_left, [aspect, 1.0, 1.0])
        assert near(cam.forward, [0, 0, -1]) and near(cam.horizontal, [-2 * aspect, 0, 0]) and near(cam.vertical, [0, 2, 0])


def _one_prim(trt, orc, kind, data):
    d = np.zeros((1, 12), np.float32); d[0, :len(data)] = data
    desc = trt.scenes.SceneDesc("one", [(0, (0.0, 0.0, 0.0), 0.0)], np.array([kind], np.uint32), d, np.zeros(1, np.uint32))
    return orc.Scene.from_desc(desc)


# ---- hittable/sphere.rs:68-100 test_sphere_hit
@pytest.mark.parametrize("use_bvh", [True, False])
def test_sphere_hit_kat(orc, trt, use_bvh):
    sc = _one_prim(trt, orc, 0, [0, 0, -1, 0.5])
    r = orc.make_ray((0, 0, 0), (0, 0, -1))
    prim, t, n = sc.trace_rays(r[:3], r[3:], 0.0, use_bvh, True)
    assert prim[0] == 0 and t[0] == 0.5 and list(n[0]) == [0, 0, 1] and list(orc.ray_at(r, t[0])) == [0, 0, -0.5]
    r = orc.make_ray((0, 0, 0), (0, math.sqrt(3.0), -3.0))
    prim, t, n = sc.trace_rays(r[:3], r[3:], 0.0, use_bvh, True)
    assert prim[0] == 0 and abs(t[0] - math.sqrt(3) / 2) < 1e-2
    assert np.linalg.norm(orc.ray_at(r, t[0]) - F([0, math.sqrt(3) / 4, -0.75])) < 1e-2
    want = F([0, math.sqrt(3), 1]); want /= np.linalg.norm(want)
    assert np.linalg.norm(n[0] - want) < 1e-2
    r = orc.make_ray((0, 0, 0), (0, -1, -1))
    prim, t, _ = sc.trace_rays(r[:3], r[3:], 0.0, use_bvh, True)
    assert prim[0] == 0xFFFFFFFF


# ---- hittable/quad.rs:67-96 test_quad_hit
@pytest.mark.parametrize("use_bvh", [True, False])
def test_quad_hit_kat(orc, trt, use_bvh):
    sc = _one_prim(trt, orc, 1, [0, 0, 0, 1, 0, 0, 0, 0, 3])
    r = orc.make_ray((0, 1, 0), (0, -1, 0))
    prim, t, n = sc.trace_rays(r[:3], r[3:], 0.0, use_bvh, True)
    assert prim[0] == 0 and t[0] == 1.0 and list(orc.ray_at(r, t[0])) == [0, 0, 0] and list(n[0]) == [0, 1, 0]   # normal flipped toward the ray
    r = orc.make_ray((0, 1, 1), (0, -1, -1))
    prim, t, n = sc.trace_rays(r[:3], r[3:], 0.0, use_bvh, True)
    assert prim[0] == 0 and abs(t[0] - math.sqrt(2)) < 1e-6 and np.all(np.abs(orc.ray_at(r, t[0])) < 1e-6) and list(n[0]) == [0, 1, 0]
    r = orc.make_ray((0, 1, 1), (0, -1, -1.1))
    prim, _, _ = sc.trace_rays(r[:3], r[3:], 0.0, use_bvh, True)
    assert prim[0] == 0xFFFFFFFF


# ---- renderer/pointgen.rs:63-107: exactly spp samples per pixel; renderer/imager.rs:72-108: mean of samples + gamma 2.2
def test_sample_bookkeeping_and_imager(orc, trt):
    desc = trt.scenes.cornell(10, 5, 3, 20)
    sc = orc.Scene.from_desc(desc); cam = orc.make_camera(**desc.camera)
    p = orc.make_params(10, 5, 3, 20, desc.background, 1)
    lin, u8, st, s64 = sc.render(cam, p, want_sum64=True)
    assert st.paths == 10 * 5 * 3
    prim, t, rad, nr = sc.render_aov(cam, p, True, True)
    assert prim.shape[0] == 150 and int(nr.sum()) == st.rays
    mean = rad.reshape(50, 3, 3).astype(np.float64).sum(1) / 3.0
    assert np.array_equal(lin.reshape(50, 3), mean.astype(np.float32))
    for v in (0.0, 1e-4, 0.25, 0.5, 1.0, 7.0, float("nan")):
        g = F(v) ** (F(1.0) / F(2.2))
        want = 0 if np.isnan(g) else int(np.clip(g, F(0.0), F(0.999)) * F(255.0))
        assert abs(orc.tonemap(v) - want) <= 1


# ---- renderer/sampler/cpu.rs:131-180 test_dummy_sampling: every sample comes back
def test_liveness(orc, trt):
    desc = trt.scenes.liveness_world()
    sc = orc.Scene.from_desc(desc)
    pts = np.zeros(100, orc.SAMPLE_POINT_DTYPE)
    for i in range(100):
        tt = F(i) / F(100)
        pts[i] = (i, 0, (F(-1.0) * (F(1.0) - tt) + F(1.0) * tt, 0.0, 0.0), (0.0, 0.0, -1.0))
    out = sc.sample_batch(pts, orc.make_params(100, 1, 1, 2, (0, 0, 0), 9))
    assert list(out["x"]) == list(range(100)) and np.all(np.isfinite(out["color"]))


# ---- Philox4x32-10 Random123 known answers
def test_philox_kat(orc):
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_keying_of_primary_ray(orc, trt):
    """Stream (seed; pixel, sample, bounce 0): u = (x + xi1)/(W-1), v = (y + xi2)/(H-1), lens sample by rejection."""
    cam = orc.make_camera(10.0, 0.0, (0, 0, 0), (0, 0, -1), (0, 1, 0), 90.0, 64, 48)   # pinhole: origin == position
    x, y, seed, sample = 17, 5, 0x1234567890abcdef, 3
    r = orc.primary_ray(cam, x, y, seed, sample)
    w = orc.philox([y * 64 + x, sample, 0, 0], [seed & 0xffffffff, seed >> 32])
    u01 = lambda b: F(b >> 9) * F(2.0 ** -23)
    u = (F(x) + u01(w[0])) / F(63); v = (F(y) + u01(w[1])) / F(47)
    ul, hz, vt = (np.array(a[:], np.float32) for a in (cam.viewport_upper_left, cam.horizontal, cam.vertical))
    d = ul + u * hz - v * vt
    d = d / np.sqrt(F(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
    assert np.allclose(r[3:], d, atol=1e-6) and list(r[:3]) == [0, 0, 0]


# ---- deterministic transcendental contract: a few ulp from the true functions, exact at octant boundaries
def test_det_math_accuracy(orc):
    u = (np.arange(0, 1 << 23, 101, dtype=np.uint32).astype(np.float32)) * F(2.0 ** -23)
    s, c, cb = orc.det_math(u)
    ts, tc, tcb = np.sin(2 * np.pi * u.astype(np.float64)), np.cos(2 * np.pi * u.astype(np.float64)), np.cbrt(u.astype(np.float64))
    assert np.abs(s - ts).max() < 2.5e-7 and np.abs(c - tc).max() < 2.5e-7
    assert np.abs(cb - tcb).max() < 2.5e-7 and cb[0] == 0.0
    s, c, _ = orc.det_math(F([0.0, 0.25, 0.5, 0.75]))
    assert list(s) == [0, 1, 0, -1] and list(c) == [1, 0, -1, 0]
    assert np.abs(s * s + c * c - 1).max() < 3e-7


# ---- materials with fixed streams (unpinned by the reference; these pin the oracle's own restatement)
def test_material_scatter(orc):
    ray = orc.make_ray((0, 1, 0), (1, -1, 0))
    n = (0.0, 2.0, 0.0)   # unnormalised outward normal, as Quad::hit passes it
    ok, out, att = orc.scatter(3, (15, 15, 15), 0.0, ray, 1.0, n)                 # light.rs:17-19
    assert not ok
    ok, out, att = orc.scatter(1, (0.8, 0.7, 0.6), 0.0, ray, 1.0, n)              # metal.rs:18-25, fuzz 0: mirror
    assert ok and np.allclose(out[3:], F([1, 1, 0]) / np.sqrt(F(2)), atol=1e-6) and np.allclose(att, [0.8, 0.7, 0.6])
    assert np.allclose(out[:3], orc.ray_at(ray, 1.0))
    ok, out5, _ = orc.scatter(1, (1, 1, 1), 5.0, ray, 1.0, n, bounce=2)           # metal.rs:13 clamps fuzz to 1
    ok, out1, _ = orc.scatter(1, (1, 1, 1), 1.0, ray, 1.0, n, bounce=2)
    assert np.array_equal(bits(out5), bits(out1))
    for b in range(1, 40):                                                         # lambertian.rs:16-22: unit direction in the normal's hemisphere
        ok, out, att = orc.scatter(0, (0.5, 0.5, 0.5), 0.0, ray, 1.0, n, bounce=b)
        assert ok and abs(np.linalg.norm(out[3:]) - 1) < 1e-6 and out[4] > -1e-6
    down = orc.make_ray((0, 1, 0), (0, -1, 0))                                     # dielectric.rs:26-46 at normal incidence
    seen = set()
    for b in range(1, 200):
        ok, out, att = orc.scatter(2, (1, 1, 1), 1.5, down, 1.0, n, bounce=b)
        assert ok and abs(abs(out[4]) - 1) < 1e-6
        seen.add(int(np.sign(out[4])))
    assert seen == {-1, 1}                                                         # refracts mostly, reflects with R0 = 4 %
    inside = orc.make_ray((0, -1, 0), (1, 0.2, 0))                                 # from inside at a grazing angle: total reflection
    ok, out, _ = orc.scatter(2, (1, 1, 1), 1.5, inside, 1.0, n)
    assert ok and out[4] < 0


# ---- BVH (bvh.rs, untested in the reference) against brute force over all primitives
@pytest.mark.parametrize("seed", [3, 4])
def test_bvh_equals_bruteforce(orc, trt, seed):
    desc = trt.scenes.random_mixed(150, seed)
    sc = orc.Scene.from_desc(desc)
    rng = np.random.default_rng(seed)
    o = rng.uniform(-14, 14, (20000, 3)).astype(np.float32)
    d = rng.normal(size=(20000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    p1, t1 = sc.trace_rays(o, d, 0.001, True)
    p2, t2 = sc.trace_rays(o, d, 0.001, False)
    assert np.array_equal(bits(t1), bits(t2))          # same closest t; the id may differ only on exact ties
    tie = p1 != p2
    assert tie.mean() < 0.05
    assert (p1 != 0xFFFFFFFF).mean() > 0.2


def test_tree_shape(orc, trt):
    for desc, nodes, depth in ((trt.scenes.cornell(), 35, 6), (trt.scenes.glass_cornell(), 67, 7)):
        ts = orc.Scene.from_desc(desc).tree_stats()
        assert ts["nodes"] == nodes == 2 * desc.n_prims - 1 and ts["max_depth"] == depth


# ---- libm mode (the reference's own call sequence) agrees statistically with the deterministic contract
def test_libm_and_det_modes_agree(orc, trt):
    desc = trt.scenes.cornell_spheres(64, 64, 64, 20)
    sc = orc.Scene.from_desc(desc); cam = orc.make_camera(**desc.camera)
    a, _, sa, _ = sc.render(cam, orc.make_params(64, 64, 64, 20, desc.background, 1, math_mode=orc.MATH_DET))
    b, _, sb, _ = sc.render(cam, orc.make_params(64, 64, 64, 20, desc.background, 1, math_mode=orc.MATH_LIBM))
    assert abs(a.mean() - b.mean()) / a.mean() < 0.05 and abs(sa.rays - sb.rays) / sa.rays < 0.01


# ---- committed golden fixtures (tests/golden/make_golden.py)
def test_oracle_reproduces_golden(orc, trt):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    g = np.load(GOLDEN)
    for name, mk in mg.CASES.items():
        d = mk(trt.scenes); c = d.camera
        sc = orc.Scene.from_desc(d); cam = orc.make_camera(**c)
        p = orc.make_params(c["width"], c["height"], d.spp, d.max_bounces, d.background, d.seed)
        prim, t, rad, nr = sc.render_aov(cam, p, True, True)
        lin, u8, st, _ = sc.render(cam, p)
        assert np.array_equal(prim, g[name + "/prim"]) and np.array_equal(bits(t), bits(g[name + "/t"]))
        assert np.array_equal(bits(rad), bits(g[name + "/rad"])) and np.array_equal(nr, g[name + "/nr"])
        assert np.array_equal(bits(lin), bits(g[name + "/lin"])) and np.array_equal(u8, g[name + "/u8"])
        assert list(g[name + "/counts"]) == [st.paths, st.rays, st.node_visits, st.sphere_tests, st.quad_tests, st.hits]
