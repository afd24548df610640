"""Named scenes and render configurations as plain data.

A SceneDesc is INPUT data only: float32/uint32 numpy arrays plus camera and renderer arguments. The product
(`tiny_raytracer_b200.Scene.from_desc`) and the test oracle (`oracle/pyoracle.py`) both consume it, so both sides see
bit-identical inputs. Nothing here computes anything about the image.

Reference sources: Cornell box = /root/reference/src/main.rs:23-125 (its two `new_box` HittableLists flattened to six
quads each in call order, because `hittable::list` does not exist in the library); test worlds =
raytracer/src/renderer/renderer.rs:88-123, hittable/quad.rs:101-138, renderer/sampler/cpu.rs:99-129.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

LAMBERTIAN, METAL, DIELECTRIC, LIGHT = 0, 1, 2, 3
SPHERE, QUAD = 0, 1


@dataclass
class SceneDesc:
    name: str
    materials: list = field(default_factory=list)            # (kind, (r, g, b), param)
    prim_type: np.ndarray = None                             # uint32 [n]
    prim_data: np.ndarray = None                             # float32 [n, 12]: sphere cx cy cz r | quad corner u v
    prim_mat: np.ndarray = None                              # uint32 [n]
    camera: dict = field(default_factory=dict)               # Camera::new arguments (camera.rs:17-26)
    spp: int = 16
    max_bounces: int = 20
    background: tuple = (0.0, 0.0, 0.0)
    seed: int = 1

    @property
    def n_prims(self) -> int:
        return int(self.prim_type.shape[0])

    def with_image(self, width: int, height: int) -> "SceneDesc":
        cam = dict(self.camera, width=int(width), height=int(height))
        return SceneDesc(self.name, self.materials, self.prim_type, self.prim_data, self.prim_mat, cam, self.spp,
                         self.max_bounces, self.background, self.seed)

    def with_render(self, spp=None, max_bounces=None, seed=None) -> "SceneDesc":
        return SceneDesc(self.name, self.materials, self.prim_type, self.prim_data, self.prim_mat, dict(self.camera),
                         self.spp if spp is None else int(spp),
                         self.max_bounces if max_bounces is None else int(max_bounces), self.background,
                         self.seed if seed is None else int(seed))


class _Builder:
    def __init__(self):
        self.materials, self.types, self.data, self.mats = [], [], [], []

    def material(self, kind, rgb, param=0.0):
        self.materials.append((int(kind), tuple(float(np.float32(c)) for c in rgb), float(np.float32(param))))
        return len(self.materials) - 1

    def sphere(self, c, r, m):
        self.types.append(SPHERE); self.data.append([c[0], c[1], c[2], r] + [0.0] * 8); self.mats.append(m)

    def quad(self, corner, u, v, m):
        self.types.append(QUAD); self.data.append(list(corner) + list(u) + list(v) + [0.0] * 3); self.mats.append(m)

    def box(self, a, b, m):
        """new_box, /root/reference/src/main.rs:89-125 (order: front(+z), right, back, left, top, bottom)."""
        a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
        mn, mx = np.minimum(a, b), np.maximum(a, b)
        dx = (mx[0] - mn[0], 0.0, 0.0); dy = (0.0, mx[1] - mn[1], 0.0); dz = (0.0, 0.0, mx[2] - mn[2])
        ndx = (-dx[0], 0.0, 0.0); ndz = (0.0, 0.0, -dz[2])
        self.quad((mn[0], mn[1], mx[2]), dx, dy, m)
        self.quad((mx[0], mn[1], mx[2]), ndz, dy, m)
        self.quad((mx[0], mn[1], mn[2]), ndx, dy, m)
        self.quad((mn[0], mn[1], mn[2]), dz, dy, m)
        self.quad((mn[0], mx[1], mx[2]), dx, ndz, m)
        self.quad((mn[0], mn[1], mn[2]), dx, dz, m)

    def done(self, name, camera, spp, max_bounces, background, seed):
        n = len(self.types)
        return SceneDesc(name, self.materials,
                         np.asarray(self.types, np.uint32).reshape(n),
                         np.asarray(self.data, np.float32).reshape(n, 12),
                         np.asarray(self.mats, np.uint32).reshape(n),
                         camera, spp, max_bounces, tuple(float(np.float32(c)) for c in background), seed)


def _camera(focus_distance, defocus_angle, position, look_at, up, vertical_fov, width, height):
    return dict(focus_distance=float(focus_distance), defocus_angle=float(defocus_angle), position=tuple(position),
                look_at=tuple(look_at), up=tuple(up), vertical_fov=float(vertical_fov), width=int(width), height=int(height))


def _cornell_walls(b: _Builder, with_boxes=True):
    """build_materials + build_objects, /root/reference/src/main.rs:29-87."""
    red = b.material(LAMBERTIAN, (0.65, 0.05, 0.05))
    white = b.material(LAMBERTIAN, (0.73, 0.73, 0.73))
    green = b.material(LAMBERTIAN, (0.12, 0.45, 0.15))
    light = b.material(LIGHT, (15.0, 15.0, 15.0))
    b.quad((100, 0, 0), (0, 100, 0), (0, 0, 100), green)
    b.quad((0, 0, 0), (0, 100, 0), (0, 0, 100), red)
    b.quad((65, 100, 60), (-30, 0, 0), (0, 0, -20), light)
    b.quad((0, 0, 0), (100, 0, 0), (0, 0, 100), white)
    b.quad((100, 100, 100), (-100, 0, 0), (0, 0, -100), white)
    b.quad((0, 0, 100), (100, 0, 0), (0, 100, 0), white)
    if with_boxes:
        b.box((25, 0, 50), (55, 60, 80), white)
        b.box((45, 0, 10), (75, 30, 40), white)
    return red, white, green, light


def _cornell_camera(width, height):
    # /root/reference/src/main.rs:8-16 with the image size replaced
    return _camera(140.0, 0.6, (50.0, 50.0, -140.0), (50.0, 50.0, 0.0), (0.0, 1.0, 0.0), 40.0, width, height)


def cornell(width=400, height=400, spp=16, max_bounces=20, seed=1) -> SceneDesc:
    """C1 (400x400x16, depth 20) and C3 (1920x1080x1024, depth 50): 18 quads, 4 materials, background 0.001."""
    b = _Builder(); _cornell_walls(b)
    return b.done("cornell", _cornell_camera(width, height), spp, max_bounces, (0.001, 0.001, 0.001), seed)


def cornell_spheres(width=400, height=400, spp=16, max_bounces=20, seed=1) -> SceneDesc:
    """C1s: the Cornell walls with a glass and a metal sphere instead of the boxes."""
    b = _Builder(); _cornell_walls(b, with_boxes=False)
    glass = b.material(DIELECTRIC, (1.0, 1.0, 1.0), 1.5)
    metal = b.material(METAL, (0.8, 0.85, 0.88), 0.05)
    b.sphere((40.0, 15.0, 30.0), 15.0, glass)
    b.sphere((65.0, 20.0, 65.0), 20.0, metal)
    return b.done("cornell_spheres", _cornell_camera(width, height), spp, max_bounces, (0.001, 0.001, 0.001), seed)


def glass_cornell(width=3840, height=2160, spp=4096, max_bounces=100, seed=1) -> SceneDesc:
    """C5: Cornell + 16 dielectric spheres (eta 1.5, r 9) on a 4x4 grid."""
    b = _Builder(); _cornell_walls(b)
    glass = b.material(DIELECTRIC, (1.0, 1.0, 1.0), 1.5)
    for i in range(4):
        for j in range(4):
            b.sphere((14.0 + 24.0 * i, 10.0 + 22.0 * ((i + j) % 3), 14.0 + 24.0 * j), 9.0, glass)
    return b.done("glass_cornell", _cornell_camera(width, height), spp, max_bounces, (0.001, 0.001, 0.001), seed)


def final_scene(width=1200, height=675, spp=500, max_bounces=50, seed=2) -> SceneDesc:
    """C2: 'Ray Tracing in One Weekend' final scene (source absent from the reference): ground sphere,
    22x22 jittered small spheres (80 % diffuse, 15 % metal, 5 % glass), three big spheres; constant background."""
    rng = np.random.Generator(np.random.Philox(key=2))
    b = _Builder()
    ground = b.material(LAMBERTIAN, (0.5, 0.5, 0.5))
    b.sphere((0.0, -1000.0, 0.0), 1000.0, ground)
    for a in range(-11, 11):
        for c in range(-11, 11):
            choose = rng.random()
            center = (a + 0.9 * rng.random(), 0.2, c + 0.9 * rng.random())
            if np.sqrt((center[0] - 4.0) ** 2 + (center[1] - 0.2) ** 2 + center[2] ** 2) <= 0.9:
                continue
            if choose < 0.8:
                m = b.material(LAMBERTIAN, tuple(rng.random(3) * rng.random(3)))
            elif choose < 0.95:
                m = b.material(METAL, tuple(0.5 + 0.5 * rng.random(3)), 0.5 * rng.random())
            else:
                m = b.material(DIELECTRIC, (1.0, 1.0, 1.0), 1.5)
            b.sphere(center, 0.2, m)
    b.sphere((0.0, 1.0, 0.0), 1.0, b.material(DIELECTRIC, (1.0, 1.0, 1.0), 1.5))
    b.sphere((-4.0, 1.0, 0.0), 1.0, b.material(LAMBERTIAN, (0.4, 0.2, 0.1)))
    b.sphere((4.0, 1.0, 0.0), 1.0, b.material(METAL, (0.7, 0.6, 0.5), 0.0))
    cam = _camera(10.0, 0.6, (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 20.0, width, height)
    return b.done("final_scene", cam, spp, max_bounces, (0.7, 0.8, 1.0), seed)


def random_spheres(n=1_000_000, width=1920, height=1080, spp=64, max_bounces=50, seed=4) -> SceneDesc:
    """C4 BVH stress: n spheres, centres uniform in [-50,50]^3, radius uniform [0.1,0.3], 80/15/5 % material mix."""
    rng = np.random.Generator(np.random.Philox(key=4))
    mats = [(LAMBERTIAN, (0.5, 0.5, 0.5), 0.0), (METAL, (0.8, 0.8, 0.8), float(np.float32(0.1))), (DIELECTRIC, (1.0, 1.0, 1.0), 1.5)]
    data = np.zeros((n, 12), np.float32)
    data[:, 0:3] = rng.uniform(-50.0, 50.0, (n, 3)).astype(np.float32)
    data[:, 3] = rng.uniform(0.1, 0.3, n).astype(np.float32)
    pick = rng.random(n)
    mat = np.where(pick < 0.8, 0, np.where(pick < 0.95, 1, 2)).astype(np.uint32)
    cam = _camera(150.0, 0.0, (0.0, 0.0, 150.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, width, height)
    return SceneDesc("random_spheres_%d" % n, mats, np.zeros(n, np.uint32), data, mat, cam, spp, max_bounces,
                     tuple(float(np.float32(c)) for c in (0.7, 0.8, 1.0)), seed)


def five_spheres(width=400, height=300, spp=3, max_bounces=10, seed=7) -> SceneDesc:
    """dummy_world + test_rendering of raytracer/src/renderer/renderer.rs:88-150 (hollow glass sphere)."""
    b = _Builder()
    ground = b.material(LAMBERTIAN, (0.0, 1.0, 0.0)); center = b.material(LAMBERTIAN, (1.0, 0.0, 0.0))
    lo = b.material(DIELECTRIC, (1.0, 1.0, 1.0), 1.5); li = b.material(DIELECTRIC, (1.0, 1.0, 1.0), np.float32(1.0) / np.float32(1.5))
    right = b.material(METAL, (0.4, 0.4, 1.0), 0.3)
    b.sphere((0.0, -100.5, -1.0), 100.0, ground); b.sphere((0.0, 0.0, -1.2), 0.5, center)
    b.sphere((1.0, 0.0, -1.0), 0.5, lo); b.sphere((1.0, 0.0, -1.0), 0.4, li); b.sphere((-1.0, 0.0, -1.0), 0.5, right)
    cam = _camera(3.4, 10.0, (-2.0, 2.0, 1.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 20.0, width, height)
    return b.done("five_spheres", cam, spp, max_bounces, (0.7, 0.8, 1.0), seed)


def five_quads(width=400, height=300, spp=10, max_bounces=10, seed=8) -> SceneDesc:
    """test_rendering of raytracer/src/hittable/quad.rs:101-151."""
    b = _Builder()
    red = b.material(LAMBERTIAN, (1.0, 0.2, 0.2)); green = b.material(LAMBERTIAN, (0.2, 1.0, 0.2))
    blue = b.material(LAMBERTIAN, (0.2, 0.2, 1.0)); orange = b.material(LAMBERTIAN, (1.0, 0.5, 0.0)); teal = b.material(LAMBERTIAN, (0.2, 0.8, 0.8))
    b.quad((-3, -2, 5), (0, 0, -4), (0, 4, 0), red); b.quad((-2, -2, 0), (4, 0, 0), (0, 4, 0), green)
    b.quad((3, -2, 1), (0, 0, 4), (0, 4, 0), blue); b.quad((-2, 3, 1), (4, 0, 0), (0, 0, 4), orange)
    b.quad((-2, -3, 5), (4, 0, 0), (0, 0, -4), teal)
    cam = _camera(1.0, 0.0, (0.0, 0.0, 9.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 80.0, width, height)
    return b.done("five_quads", cam, spp, max_bounces, (0.7, 0.8, 1.0), seed)


def liveness_world(seed=9) -> SceneDesc:
    """dummy_world of raytracer/src/renderer/sampler/cpu.rs:99-129 (five white Lambertian spheres)."""
    b = _Builder(); m = b.material(LAMBERTIAN, (1.0, 1.0, 1.0))
    b.sphere((0.0, -100.5, -1.0), 100.0, m); b.sphere((0.0, 0.0, -1.2), 0.5, m); b.sphere((1.0, 0.0, -1.0), 0.5, m)
    b.sphere((1.0, 0.0, -1.0), 0.4, m); b.sphere((-1.0, 0.0, -1.0), 0.5, m)
    cam = _camera(1.0, 0.0, (0.0, 0.0, 0.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), 90.0, 100, 1)
    return b.done("liveness", cam, 1, 2, (0.0, 0.0, 0.0), seed)


def random_mixed(n=200, seed=11, width=64, height=48, spp=2, max_bounces=12) -> SceneDesc:
    """Parity stress world: spheres and quads of all four materials at random, including overlapping, coplanar and
    duplicate primitives so that exact-t ties and the reference's leftmost-leaf rule are exercised."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    b = _Builder()
    mats = [b.material(LAMBERTIAN, tuple(rng.random(3))), b.material(METAL, tuple(rng.random(3)), rng.random()),
            b.material(METAL, (0.9, 0.9, 0.9), 0.0), b.material(DIELECTRIC, (1.0, 1.0, 1.0), 1.5),
            b.material(DIELECTRIC, (0.9, 1.0, 0.9), 1.2), b.material(LIGHT, (4.0, 3.0, 2.0)), b.material(LAMBERTIAN, (0.8, 0.8, 0.8))]
    for i in range(n):
        m = mats[int(rng.integers(len(mats)))]
        kind = rng.random()
        if kind < 0.5:
            b.sphere(tuple(rng.uniform(-10, 10, 3)), float(rng.uniform(0.3, 2.5)), m)
        elif kind < 0.8:
            c = rng.uniform(-10, 10, 3); u = rng.uniform(-4, 4, 3); v = rng.uniform(-4, 4, 3)
            b.quad(tuple(c), tuple(u), tuple(v), m)
        elif kind < 0.9:
            # axis-aligned quad on an integer plane: coplanar overlaps between different primitives
            z = float(rng.integers(-3, 4)); c = np.round(rng.uniform(-8, 8, 2))
            b.quad((c[0], c[1], z), (float(rng.integers(2, 6)), 0.0, 0.0), (0.0, float(rng.integers(2, 6)), 0.0), m)
        elif len(b.types):
            # exact duplicate of an earlier primitive with another material
            j = int(rng.integers(len(b.types)))
            b.types.append(b.types[j]); b.data.append(list(b.data[j])); b.mats.append(m)
    cam = _camera(20.0, 0.5, (0.0, 2.0, 24.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0, width, height)
    return b.done("random_mixed_%d_%d" % (n, seed), cam, spp, max_bounces, (0.6, 0.7, 0.9), seed)


WORKLOADS = {
    # BASELINE.json configs: name -> factory at full size
    "c1": lambda: cornell(400, 400, 16, 20),
    "c1s": lambda: cornell_spheres(400, 400, 16, 20),
    "c2": lambda: final_scene(1200, 675, 500, 50),
    "c3": lambda: cornell(1920, 1080, 1024, 50),
    "c4": lambda: random_spheres(1_000_000, 1920, 1080, 64, 50),
    "c5": lambda: glass_cornell(3840, 2160, 4096, 100),
}
