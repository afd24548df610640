"""Host-side mirror of tiny-raytracer's public API over the C ABI (include/trt.h).

Two layers:
  * `Context`, `Scene`: thin handles over trt_ctx / trt_scene (what a Rust `trt-sys` crate would wrap).
  * `Vec3`, `Lambertian`/`Metal`/`Dielectric`/`Light`, `Sphere`/`Quad`, `World`, `Camera`, `Renderer`, `Image`:
    the names, argument meanings and error behaviour of the reference's
    `raytracer::{math::vec3::Vec3, material::*, hittable::{sphere::Sphere, quad::Quad, world::World}, camera::Camera,
    renderer::Renderer, utils::image::Image}` (raytracer/src/lib.rs:8-14), so reference scene scripts such as
    /root/reference/src/main.rs translate line by line.
Everything that computes pixels runs in libtrt_b200.so on the GPU; this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import struct
import zlib

import numpy as np

from . import _lib as L
from .scenes import SceneDesc, LAMBERTIAN, METAL, DIELECTRIC, LIGHT, SPHERE, QUAD

BVH_REFERENCE, BVH_SAH, BVH_FLAT, BVH_AUTO, BVH_LBVH_DEVICE = L.BVH_REFERENCE, L.BVH_SAH, L.BVH_FLAT, L.BVH_AUTO, L.BVH_LBVH_DEVICE
VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT = L.VARIANT_MEGAKERNEL, L.VARIANT_WAVEFRONT


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------- low level handles
class Context:
    """One GPU (trt_ctx). There is no CPU path: construction raises TrtError when no B200-class device is usable."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.trt_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self.stream_handle = None      # None = the ctx's own non-blocking stream; else the cudaStream_t given to set_stream

    def close(self):
        if getattr(self, "_h", None):
            self._lib.trt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def name(self) -> str:
        buf = C.create_string_buffer(256)
        L.check(self._lib.trt_ctx_device_name(self._h, buf, 256))
        return buf.value.decode()

    @property
    def sm_count(self) -> int:
        n = C.c_int()
        L.check(self._lib.trt_ctx_sm_count(self._h, C.byref(n)))
        return n.value

    def synchronize(self):
        L.check(self._lib.trt_ctx_synchronize(self._h))

    def set_stream(self, cuda_stream: int):
        """Run on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = legacy default stream."""
        L.check(self._lib.trt_ctx_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))
        self.stream_handle = int(cuda_stream or 0)


class MultiContext:
    """Several GPUs of one box behind one handle (trt_multi): contexts, peer mappings and buffers are created once and
    reused across renders. `devices` may list a device more than once (two shares on one GPU: tests)."""

    def __init__(self, devices):
        self._lib = L.load()
        devs = [int(d) for d in devices]
        arr = (C.c_int * max(len(devs), 1))(*devs)
        h = C.c_void_p()
        L.check(self._lib.trt_multi_create(arr, len(devs), C.byref(h)))
        self._h, self.devices = h, devs

    def close(self):
        if getattr(self, "_h", None):
            self._lib.trt_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def peer_direct(self) -> bool:
        n, d = C.c_int(), C.c_int()
        L.check(self._lib.trt_multi_device_count(self._h, C.byref(n), C.byref(d)))
        return bool(d.value)

    def render(self, scene: "Scene", cam: L.Camera, params: L.Params, want_linear=True, want_rgb8=True, out_linear=None, out_rgb8=None):
        """trt_multi_render: HOST buffers out; with both outputs off the result stays on the devices (timing)."""
        n = params.width * params.height * 3
        lin = (out_linear if out_linear is not None else np.empty(n, np.float32)) if want_linear else None
        u8 = (out_rgb8 if out_rgb8 is not None else np.empty(n, np.uint8)) if want_rgb8 else None
        st = L.Stats()
        L.check(self._lib.trt_multi_render(self._h, scene._h, C.byref(cam), C.byref(params), _ptr(lin), _ptr(u8), C.byref(st)))
        shape = (params.height, params.width, 3)
        return (None if lin is None else lin.reshape(shape)), (None if u8 is None else u8.reshape(shape)), st


def make_camera(focus_distance, defocus_angle, position, look_at, up, vertical_fov, width, height) -> L.Camera:
    """Camera::new (camera.rs:17-56) -> plain-data camera."""
    cam = L.Camera()
    L.check(L.load().trt_camera_make(float(focus_distance), float(defocus_angle), L.f3(position), L.f3(look_at), L.f3(up),
                                     float(vertical_fov), int(width), int(height), C.byref(cam)))
    return cam


def make_params(width, height, spp, max_bounces, background=(0, 0, 0), seed=1, sample_offset=0, spp_total=0,
                t_min=0.001, variant=VARIANT_MEGAKERNEL, flags=0) -> L.Params:
    p = L.Params()
    p.width, p.height, p.spp, p.sample_offset, p.max_bounces = int(width), int(height), int(spp), int(sample_offset), int(max_bounces)
    p.background[0], p.background[1], p.background[2] = [float(c) for c in background]
    p.seed, p.t_min, p.spp_total, p.variant, p.flags = int(seed), float(t_min), int(spp_total), int(variant), int(flags)
    return p


class Scene:
    """trt_scene: materials + primitives, committed to a BVH of the chosen kind."""

    def __init__(self):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.trt_scene_create(C.byref(h)))
        self._h = h
        self.committed_kind = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.trt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_material(self, kind, rgb, param=0.0) -> int:
        mid = C.c_uint32()
        if kind == LAMBERTIAN:
            L.check(self._lib.trt_add_lambertian(self._h, L.f3(rgb), C.byref(mid)))
        elif kind == METAL:
            L.check(self._lib.trt_add_metal(self._h, L.f3(rgb), float(param), C.byref(mid)))
        elif kind == DIELECTRIC:
            L.check(self._lib.trt_add_dielectric(self._h, L.f3(rgb), float(param), C.byref(mid)))
        elif kind == LIGHT:
            L.check(self._lib.trt_add_light(self._h, L.f3(rgb), C.byref(mid)))
        else:
            raise ValueError("unknown material kind %r" % (kind,))
        return mid.value

    def add_sphere(self, center, radius, mat) -> int:
        pid = C.c_uint32()
        L.check(self._lib.trt_add_sphere(self._h, L.f3(center), float(radius), int(mat), C.byref(pid)))
        return pid.value

    def add_quad(self, corner, u, v, mat) -> int:
        pid = C.c_uint32()
        L.check(self._lib.trt_add_quad(self._h, L.f3(corner), L.f3(u), L.f3(v), int(mat), C.byref(pid)))
        return pid.value

    def add_prims(self, prim_type, prim_data, prim_mat):
        t = np.ascontiguousarray(prim_type, np.uint32); d = np.ascontiguousarray(prim_data, np.float32).reshape(-1, 12)
        m = np.ascontiguousarray(prim_mat, np.uint32)
        assert t.shape[0] == d.shape[0] == m.shape[0]
        L.check(self._lib.trt_add_prims(self._h, int(t.shape[0]), _ptr(t), _ptr(d), _ptr(m)))

    @classmethod
    def from_desc(cls, desc: SceneDesc, bvh_kind=BVH_SAH) -> "Scene":
        s = cls()
        for kind, rgb, param in desc.materials:
            s.add_material(kind, rgb, param)
        s.add_prims(desc.prim_type, desc.prim_data, desc.prim_mat)
        s.commit(bvh_kind)
        return s

    @property
    def num_prims(self) -> int:
        n = C.c_uint32()
        L.check(self._lib.trt_scene_num_prims(self._h, C.byref(n)))
        return n.value

    def commit(self, bvh_kind=BVH_SAH):
        L.check(self._lib.trt_scene_commit(self._h, int(bvh_kind)))
        self.committed_kind = int(bvh_kind)

    def leaf_ranks(self) -> np.ndarray:
        r = np.zeros(self.num_prims, np.uint32)
        L.check(self._lib.trt_scene_leaf_ranks(self._h, _ptr(r)))
        return r

    def bvh_info(self):
        n, d, c = C.c_uint32(), C.c_uint32(), C.c_float()
        L.check(self._lib.trt_scene_bvh_info(self._h, C.byref(n), C.byref(d), C.byref(c)))
        return {"nodes": n.value, "max_depth": d.value, "sah_cost": c.value}

    def upload(self, ctx: Context):
        L.check(self._lib.trt_scene_upload(ctx._h, self._h))

    def build_times(self):
        h, d = C.c_float(), C.c_float()
        L.check(self._lib.trt_scene_build_times(self._h, C.byref(h), C.byref(d)))
        return {"host_commit_ms": h.value, "device_build_ms": d.value}

    # ---- rendering entry points (HOST buffers)
    def render(self, ctx: Context, cam: L.Camera, params: L.Params, want_linear=True, want_rgb8=True, out_linear=None, out_rgb8=None):
        """trt_render: HOST buffers out. `out_linear` / `out_rgb8` let the caller supply the buffers (flat float32 / uint8
        arrays of H*W*3 values, e.g. views of pinned memory) instead of fresh numpy arrays."""
        n = params.width * params.height * 3
        lin = (out_linear if out_linear is not None else np.empty(n, np.float32)) if want_linear else None
        u8 = (out_rgb8 if out_rgb8 is not None else np.empty(n, np.uint8)) if want_rgb8 else None
        assert lin is None or (lin.dtype == np.float32 and lin.size == n and lin.flags.c_contiguous)
        assert u8 is None or (u8.dtype == np.uint8 and u8.size == n and u8.flags.c_contiguous)
        st = L.Stats()
        L.check(self._lib.trt_render(ctx._h, self._h, C.byref(cam), C.byref(params), _ptr(lin), _ptr(u8), C.byref(st)))
        shape = (params.height, params.width, 3)
        return (None if lin is None else lin.reshape(shape)), (None if u8 is None else u8.reshape(shape)), st

    def render_multi(self, devices, cam: L.Camera, params: L.Params):
        """trt_render_multi: one call, several GPUs (create + render + destroy); HOST buffers out."""
        devs = [int(d) for d in devices]
        arr = (C.c_int * max(len(devs), 1))(*devs)
        n = params.width * params.height * 3
        lin = np.empty(n, np.float32); u8 = np.empty(n, np.uint8)
        st = L.Stats()
        L.check(self._lib.trt_render_multi(arr, len(devs), self._h, C.byref(cam), C.byref(params), _ptr(lin), _ptr(u8), C.byref(st)))
        shape = (params.height, params.width, 3)
        return lin.reshape(shape), u8.reshape(shape), st

    def render_accumulate(self, ctx: Context, cam: L.Camera, params: L.Params, d_sum64_ptr: int, want_stats=True):
        st = L.Stats() if want_stats else None
        L.check(self._lib.trt_render_accumulate(ctx._h, self._h, C.byref(cam), C.byref(params), C.c_void_p(d_sum64_ptr),
                                                C.byref(st) if st is not None else None))
        return st

    def render_aov(self, ctx: Context, cam: L.Camera, params: L.Params, want_radiance=False, want_n_rays=False):
        n = params.width * params.height * params.spp
        prim = np.empty(n, np.uint32); t = np.empty(n, np.float32)
        rad = np.empty((n, 3), np.float32) if want_radiance else None
        nr = np.empty(n, np.uint32) if want_n_rays else None
        L.check(self._lib.trt_render_aov(ctx._h, self._h, C.byref(cam), C.byref(params), _ptr(prim), _ptr(t), _ptr(rad), _ptr(nr)))
        return prim, t, rad, nr

    def sample_batch(self, ctx: Context, points: np.ndarray, params: L.Params) -> np.ndarray:
        """points: structured array with the SamplePoint layout (see SAMPLE_POINT_DTYPE); returns SAMPLED_COLOR_DTYPE."""
        pts = np.ascontiguousarray(points, SAMPLE_POINT_DTYPE)
        out = np.empty(pts.shape[0], SAMPLED_COLOR_DTYPE)
        L.check(self._lib.trt_sample_batch(ctx._h, self._h, _ptr(pts), int(pts.shape[0]), C.byref(params), _ptr(out)))
        return out

    def trace_rays(self, ctx: Context, origins, dirs, t_min=0.001):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3); d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        prim = np.empty(n, np.uint32); t = np.empty(n, np.float32)
        L.check(self._lib.trt_trace_rays(ctx._h, self._h, n, _ptr(o), _ptr(d), float(t_min), _ptr(prim), _ptr(t)))
        return prim, t


class SamplerStream:
    """Streaming `Sampler` adapter (trt_sampler_*): `sampling(batches)` consumes an iterable of SamplePoint arrays and
    yields the SampledColor arrays in order, keeping two batches in flight (copy/compute overlap on two streams) — the
    shape of `Sampler::sampling(world, in_channel, out_channel)` (renderer/sampler/mod.rs:9-16) with channels replaced
    by an iterator. The adapter ends when the input ends, as the reference's samplers do when `in_channel` closes."""

    def __init__(self, ctx: Context, scene: "Scene", params: L.Params, batch_capacity: int = 10240):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.trt_sampler_create(ctx._h, scene._h, C.byref(params), int(batch_capacity), C.byref(h)))
        self._h, self.capacity = h, int(batch_capacity)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.trt_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, points: np.ndarray):
        pts = np.ascontiguousarray(points, SAMPLE_POINT_DTYPE)
        L.check(self._lib.trt_sampler_submit(self._h, _ptr(pts), int(pts.shape[0])))

    def collect(self):
        out = np.empty(self.capacity, SAMPLED_COLOR_DTYPE)
        n = C.c_size_t()
        L.check(self._lib.trt_sampler_collect(self._h, _ptr(out), C.byref(n)))
        return out[:n.value]

    def sampling(self, batches):
        in_flight = 0
        for b in batches:
            for lo in range(0, max(len(b), 1), self.capacity):
                if in_flight == 2:
                    yield self.collect(); in_flight -= 1
                self.submit(b[lo:lo + self.capacity]); in_flight += 1
        while in_flight:
            yield self.collect(); in_flight -= 1


SAMPLE_POINT_DTYPE = np.dtype([("x", np.uint32), ("y", np.uint32), ("origin", np.float32, 3), ("direction", np.float32, 3)])
SAMPLED_COLOR_DTYPE = np.dtype([("x", np.uint32), ("y", np.uint32), ("color", np.float32, 3)])
assert SAMPLE_POINT_DTYPE.itemsize == 32 and SAMPLED_COLOR_DTYPE.itemsize == 20


def finalize(ctx: Context, d_sum64_ptr: int, width: int, height: int, spp_total: int, d_linear_ptr: int = 0, d_rgb8_ptr: int = 0):
    """trt_finalize on DEVICE pointers."""
    L.check(L.load().trt_finalize(ctx._h, C.c_void_p(d_sum64_ptr), int(width), int(height), int(spp_total),
                                  C.c_void_p(d_linear_ptr) if d_linear_ptr else None,
                                  C.c_void_p(d_rgb8_ptr) if d_rgb8_ptr else None, None))


def debug_math(ctx: Context, u: np.ndarray):
    u = np.ascontiguousarray(u, np.float32)
    s, c, cb = np.empty_like(u), np.empty_like(u), np.empty_like(u)
    L.check(L.load().trt_debug_math(ctx._h, int(u.shape[0]), _ptr(u), _ptr(s), _ptr(c), _ptr(cb)))
    return s, c, cb


def debug_philox(ctx: Context, ctr: np.ndarray, key) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, np.uint32).reshape(-1, 4)
    out = np.empty_like(ctr)
    k = (C.c_uint32 * 2)(int(key[0]), int(key[1]))
    L.check(L.load().trt_debug_philox(ctx._h, int(ctr.shape[0]), _ptr(ctr), k, _ptr(out)))
    return out


# ----------------------------------------------------------------------------- the reference's API surface
class Vec3:
    """math::vec3::Vec3 (math/vec3.rs:9-30): three f32."""
    __slots__ = ("x", "y", "z")

    def __init__(self, x, y, z):
        self.x, self.y, self.z = float(np.float32(x)), float(np.float32(y)), float(np.float32(z))

    @classmethod
    def new(cls, x, y, z): return cls(x, y, z)

    @classmethod
    def new_diagonal(cls, v): return cls(v, v, v)

    @classmethod
    def zero(cls): return cls(0.0, 0.0, 0.0)

    def __iter__(self): return iter((self.x, self.y, self.z))

    def __getitem__(self, i): return (self.x, self.y, self.z)[i]

    def __neg__(self): return Vec3(-self.x, -self.y, -self.z)

    def __repr__(self): return "(%g, %g, %g)" % (self.x, self.y, self.z)


def _v(v):
    return tuple(float(np.float32(c)) for c in v)


class Material:
    kind = None


class Lambertian(Material):
    kind = LAMBERTIAN

    def __init__(self, albedo): self.rgb, self.param = _v(albedo), 0.0


class Metal(Material):
    kind = METAL

    def __init__(self, albedo, fuzz): self.rgb, self.param = _v(albedo), float(fuzz)


class Dielectric(Material):
    kind = DIELECTRIC

    def __init__(self, albedo, refraction_index): self.rgb, self.param = _v(albedo), float(refraction_index)


class Light(Material):
    kind = LIGHT

    def __init__(self, color): self.rgb, self.param = _v(color), 0.0


class Sphere:
    """hittable::sphere::Sphere::new(center, radius, material) (hittable/sphere.rs:16-25)."""

    def __init__(self, center, radius, material: Material):
        self.center, self.radius, self.material = _v(center), float(radius), material


class Quad:
    """hittable::quad::Quad::new(corner, u, v, material) (hittable/quad.rs:20-29)."""

    def __init__(self, corner, u, v, material: Material):
        self.corner, self.u, self.v, self.material = _v(corner), _v(u), _v(v), material


def new_box(a, b, mat: Material):
    """new_box of /root/reference/src/main.rs:89-125; returns the six quads (the reference's HittableList is flattened)."""
    a, b = np.asarray(_v(a), np.float32), np.asarray(_v(b), np.float32)
    mn, mx = np.minimum(a, b), np.maximum(a, b)
    dx = (mx[0] - mn[0], 0.0, 0.0); dy = (0.0, mx[1] - mn[1], 0.0); dz = (0.0, 0.0, mx[2] - mn[2])
    ndx = (-dx[0], 0.0, 0.0); ndz = (0.0, 0.0, -dz[2])
    return [Quad((mn[0], mn[1], mx[2]), dx, dy, mat), Quad((mx[0], mn[1], mx[2]), ndz, dy, mat),
            Quad((mx[0], mn[1], mn[2]), ndx, dy, mat), Quad((mn[0], mn[1], mn[2]), dz, dy, mat),
            Quad((mn[0], mx[1], mx[2]), dx, ndz, mat), Quad((mn[0], mn[1], mn[2]), dx, dz, mat)]


class Transform:
    """math::transform::Transform (math/transform.rs:10-111): a 4x4 f32 matrix; `new(translation, rotation, scale)` =
    translate * scale * rotate with rotate = Rz * Ry * Rx (transform.rs:15-21,78-80); `t * v` transforms a point (:104-111).
    The reference never applies it to geometry (dead code upstream). Here it also instances geometry:
    `World.add_instance(geometries, transform)` adds transformed COPIES — a quad's corner goes through the matrix, its edges
    through the linear part; a sphere needs a similarity (rotation + uniform |scale|) — so one template (e.g. `new_box`)
    can be placed many times. Instances are flattened on the host like `new_box` is: the device sees ordinary primitives."""

    def __init__(self, mat=None):
        self.mat = np.eye(4, dtype=np.float32) if mat is None else np.asarray(mat, np.float32).reshape(4, 4)

    @classmethod
    def new(cls, translation, rotation, scale):
        return cls.translate(translation) * cls.scale(scale) * cls.rotate(rotation)

    @classmethod
    def identity(cls): return cls()

    @classmethod
    def translate(cls, delta):
        m = np.eye(4, dtype=np.float32); m[0, 3], m[1, 3], m[2, 3] = _v(delta); return cls(m)

    @classmethod
    def scale(cls, scales):
        m = np.eye(4, dtype=np.float32); m[0, 0], m[1, 1], m[2, 2] = _v(scales); return cls(m)

    @staticmethod
    def _cs(angle):
        a = np.float32(angle)
        return np.float32(np.cos(a)), np.float32(np.sin(a))

    @classmethod
    def rotate_x(cls, angle):
        c, s = cls._cs(angle); m = np.eye(4, dtype=np.float32); m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c; return cls(m)

    @classmethod
    def rotate_y(cls, angle):
        c, s = cls._cs(angle); m = np.eye(4, dtype=np.float32); m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c; return cls(m)

    @classmethod
    def rotate_z(cls, angle):
        c, s = cls._cs(angle); m = np.eye(4, dtype=np.float32); m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c; return cls(m)

    @classmethod
    def rotate(cls, angles):
        ax, ay, az = _v(angles)
        return cls.rotate_z(az) * cls.rotate_y(ay) * cls.rotate_x(ax)

    def __mul__(self, rhs):
        if isinstance(rhs, Transform):
            return Transform(self.mat @ rhs.mat)
        x, y, z = _v(rhs)
        r = self.mat @ np.array([x, y, z, 1.0], np.float32)
        return Vec3(r[0], r[1], r[2])

    def linear(self, v):
        """The matrix without its translation column applied to a direction / edge vector."""
        x, y, z = _v(v)
        r = self.mat[:3, :3] @ np.array([x, y, z], np.float32)
        return Vec3(r[0], r[1], r[2])

    def apply(self, geometry):
        """A transformed copy of a Sphere / Quad (or list of them)."""
        if isinstance(geometry, (list, tuple)):
            return [self.apply(g) for g in geometry]
        if isinstance(geometry, Quad):
            return Quad(self * geometry.corner, self.linear(geometry.u), self.linear(geometry.v), geometry.material)
        if isinstance(geometry, Sphere):
            L3 = self.mat[:3, :3].astype(np.float64)
            g = L3.T @ L3                                            # similarity <=> L^T L = k^2 I
            k2 = float(g[0, 0])
            if k2 <= 0.0 or not np.allclose(g, k2 * np.eye(3), rtol=1e-5, atol=1e-6 * k2):
                raise ValueError("Transform.apply: a sphere needs a similarity transform (rotation + uniform scale)")
            return Sphere(self * geometry.center, float(np.float32(geometry.radius) * np.float32(np.sqrt(k2))), geometry.material)
        raise TypeError("unsupported geometry %r" % (geometry,))


class World:
    """hittable::world::World (hittable/world.rs:10-55)."""

    def __init__(self):
        self.geometries, self.materials = [], {}

    @classmethod
    def new(cls): return cls()

    def add_geometry(self, geometry):
        if isinstance(geometry, (list, tuple)):     # a flattened HittableList (new_box)
            self.geometries.extend(geometry)
        else:
            self.geometries.append(geometry)

    def add_instance(self, geometry, transform: "Transform"):
        """Instanced geometry: a transformed copy of `geometry` (one Sphere / Quad or a list such as `new_box(...)`)."""
        self.add_geometry(transform.apply(geometry))

    def add_material(self, name: str, material: Material):
        if name in self.materials:                  # world.rs:29-31 panics
            raise KeyError("%s key is already in the material table" % name)
        self.materials[name] = material

    def get_material(self, name: str):
        return self.materials.get(name)

    def get_geometries(self, cls):
        return [g for g in self.geometries if isinstance(g, cls)]

    def to_desc(self, name="world") -> SceneDesc:
        mats, index = [], {}
        def mid(m):
            if id(m) not in index:
                index[id(m)] = len(mats); mats.append((m.kind, m.rgb, m.param))
            return index[id(m)]
        for m in self.materials.values():
            mid(m)
        n = len(self.geometries)
        t = np.zeros(n, np.uint32); d = np.zeros((n, 12), np.float32); pm = np.zeros(n, np.uint32)
        for i, g in enumerate(self.geometries):
            if isinstance(g, Sphere):
                t[i] = SPHERE; d[i, 0:3] = g.center; d[i, 3] = g.radius
            elif isinstance(g, Quad):
                t[i] = QUAD; d[i, 0:3] = g.corner; d[i, 3:6] = g.u; d[i, 6:9] = g.v
            else:
                raise TypeError("unsupported geometry %r" % (g,))
            pm[i] = mid(g.material)
        return SceneDesc(name, mats, t, d, pm)


class Camera:
    """camera::Camera::new(focus_distance, defocus_angle, position, look_at, up, vertical_fov, width, height) (camera.rs:17-56)."""

    def __init__(self, focus_distance, defocus_angle, position, look_at, up, vertical_fov, width, height):
        self.pod = make_camera(focus_distance, defocus_angle, _v(position), _v(look_at), _v(up), vertical_fov, width, height)

    @classmethod
    def new(cls, *a): return cls(*a)

    def get_image_size(self): return (self.pod.width, self.pod.height)


class Image:
    """utils::image::Image (utils/image.rs): linear f32 pixels plus the 8-bit tone-mapped form, `save` writes a PNG."""

    def __init__(self, linear: np.ndarray, rgb8: np.ndarray):
        self.linear, self.rgb8 = linear, rgb8

    def width(self): return self.rgb8.shape[1]

    def height(self): return self.rgb8.shape[0]

    def size(self): return (self.width(), self.height())

    def get_pixel(self, x, y): return tuple(self.linear[y, x])

    def save(self, filename: str):
        h, w, _ = self.rgb8.shape
        raw = b"".join(b"\x00" + self.rgb8[y].tobytes() for y in range(h))
        def chunk(tag, data):
            c = struct.pack(">I", len(data)) + tag + data
            return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
        png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + \
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
        with open(filename, "wb") as f:
            f.write(png)


class Renderer:
    """renderer::Renderer::new(samples_per_pixel, num_sampler_threads, max_bounces, progressbar, background_color)
    and `render(camera, world) -> Image` (renderer/renderer.rs:21-79). `num_sampler_threads` and `progressbar` are
    accepted for signature compatibility and ignored: generator, sampler and imager all run on the GPU.
    `devices=[0, 1, ...]` renders on several GPUs of the box through trt_render_multi (samples-per-pixel partitioned,
    NVLink peer reduce + tone map on the GPUs); the default is one GPU (`device`)."""

    def __init__(self, samples_per_pixel, num_sampler_threads=0, max_bounces=20, progressbar=False, background_color=None,
                 *, seed=1, device=0, bvh_kind=BVH_AUTO, variant=VARIANT_MEGAKERNEL, devices=None):
        self.samples_per_pixel, self.max_bounces = int(samples_per_pixel), int(max_bounces)
        self.background_color = _v(background_color) if background_color is not None else (0.0, 0.0, 0.0)   # renderer.rs:33
        self.seed, self.device, self.bvh_kind, self.variant = int(seed), int(device), int(bvh_kind), int(variant)
        self.devices = [int(d) for d in devices] if devices else None
        self.last_stats = None

    @classmethod
    def new(cls, *a, **k): return cls(*a, **k)

    def render(self, camera: Camera, world: World) -> Image:
        if self.devices and len(self.devices) > 1:
            scene = Scene.from_desc(world.to_desc(), self.bvh_kind)
            w, h = camera.get_image_size()
            p = make_params(w, h, self.samples_per_pixel, self.max_bounces, self.background_color, self.seed, variant=self.variant)
            lin, u8, st = scene.render_multi(self.devices, camera.pod, p)
            self.last_stats = st
            scene.close()
            return Image(lin, u8)
        ctx = Context(self.devices[0] if self.devices else self.device)
        try:
            scene = Scene.from_desc(world.to_desc(), self.bvh_kind)
            w, h = camera.get_image_size()
            p = make_params(w, h, self.samples_per_pixel, self.max_bounces, self.background_color, self.seed, variant=self.variant)
            lin, u8, st = scene.render(ctx, camera.pod, p)
            self.last_stats = st
            scene.close()
            return Image(lin, u8)
        finally:
            ctx.close()
