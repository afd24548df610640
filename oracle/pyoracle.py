"""ctypes wrapper of oracle/liboracle.so — the CPU ORACLE (test infrastructure, see oracle/oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It consumes the same SceneDesc input data as the product but shares no code with it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("viewport_upper_left", C.c_float * 3), ("forward", C.c_float * 3),
                ("horizontal", C.c_float * 3), ("vertical", C.c_float * 3), ("defocus_disk_u", C.c_float * 3),
                ("defocus_disk_v", C.c_float * 3), ("width", C.c_uint32), ("height", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("sample_offset", C.c_uint32),
                ("max_bounces", C.c_uint32), ("background", C.c_float * 3), ("seed", C.c_uint64), ("t_min", C.c_float),
                ("spp_total", C.c_uint32), ("math_mode", C.c_uint32), ("n_threads", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("node_visits", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("quad_tests", C.c_uint64), ("hits", C.c_uint64), ("wall_ms", C.c_double), ("threads_used", C.c_uint32)]


SAMPLE_POINT_DTYPE = np.dtype([("x", np.uint32), ("y", np.uint32), ("origin", np.float32, 3), ("direction", np.float32, 3)])
SAMPLED_COLOR_DTYPE = np.dtype([("x", np.uint32), ("y", np.uint32), ("color", np.float32, 3)])

MATH_DET, MATH_LIBM = 0, 1
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.orc_scene_create.restype = C.c_void_p
        for name in ("orc_add_lambertian", "orc_add_metal", "orc_add_dielectric", "orc_add_light", "orc_add_sphere",
                     "orc_add_quad", "orc_num_prims"):
            getattr(lib, name).restype = C.c_uint32
        lib.orc_tonemap.restype = C.c_uint8
        lib.orc_tonemap.argtypes = [C.c_float]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


def make_camera(focus_distance, defocus_angle, position, look_at, up, vertical_fov, width, height) -> Camera:
    cam = Camera()
    load().orc_camera_make(C.c_float(focus_distance), C.c_float(defocus_angle), _f3(position), _f3(look_at), _f3(up),
                           C.c_float(vertical_fov), C.c_uint32(width), C.c_uint32(height), C.byref(cam))
    return cam


def make_params(width, height, spp, max_bounces, background=(0, 0, 0), seed=1, sample_offset=0, spp_total=0, t_min=0.001,
                math_mode=MATH_DET, n_threads=0) -> Params:
    p = Params()
    p.width, p.height, p.spp, p.sample_offset, p.max_bounces = int(width), int(height), int(spp), int(sample_offset), int(max_bounces)
    p.background[0], p.background[1], p.background[2] = [float(c) for c in background]
    p.seed, p.t_min, p.spp_total, p.math_mode, p.n_threads = int(seed), float(t_min), int(spp_total), int(math_mode), int(n_threads)
    return p


class Scene:
    def __init__(self):
        self._lib = load()
        self._h = C.c_void_p(self._lib.orc_scene_create())

    def __del__(self):
        try:
            if self._h:
                self._lib.orc_scene_destroy(self._h); self._h = None
        except Exception:
            pass

    @classmethod
    def from_desc(cls, desc) -> "Scene":
        s = cls(); lib = s._lib
        for kind, rgb, param in desc.materials:
            if kind == 0: lib.orc_add_lambertian(s._h, _f3(rgb))
            elif kind == 1: lib.orc_add_metal(s._h, _f3(rgb), C.c_float(param))
            elif kind == 2: lib.orc_add_dielectric(s._h, _f3(rgb), C.c_float(param))
            else: lib.orc_add_light(s._h, _f3(rgb))
        t = np.ascontiguousarray(desc.prim_type, np.uint32); d = np.ascontiguousarray(desc.prim_data, np.float32)
        m = np.ascontiguousarray(desc.prim_mat, np.uint32)
        lib.orc_add_prims(s._h, C.c_uint32(t.shape[0]), _p(t), _p(d), _p(m))
        lib.orc_scene_commit(s._h)
        return s

    @property
    def num_prims(self):
        return int(self._lib.orc_num_prims(self._h))

    def leaf_ranks(self):
        r = np.zeros(self.num_prims, np.uint32); self._lib.orc_leaf_ranks(self._h, _p(r)); return r

    def tree_stats(self):
        n, d = C.c_uint32(), C.c_uint32(); self._lib.orc_tree_stats(self._h, C.byref(n), C.byref(d))
        return {"nodes": n.value, "max_depth": d.value}

    def render(self, cam: Camera, p: Params, want_sum64=False):
        n = p.width * p.height * 3
        s64 = np.empty(n, np.float64) if want_sum64 else None
        lin = np.empty(n, np.float32); u8 = np.empty(n, np.uint8); st = Stats()
        self._lib.orc_render(self._h, C.byref(cam), C.byref(p), _p(s64), _p(lin), _p(u8), C.byref(st))
        shape = (p.height, p.width, 3)
        return lin.reshape(shape), u8.reshape(shape), st, (None if s64 is None else s64.reshape(shape))

    def render_aov(self, cam: Camera, p: Params, want_radiance=False, want_n_rays=False):
        n = p.width * p.height * p.spp
        prim = np.empty(n, np.uint32); t = np.empty(n, np.float32)
        rad = np.empty((n, 3), np.float32) if want_radiance else None
        nr = np.empty(n, np.uint32) if want_n_rays else None
        self._lib.orc_render_aov(self._h, C.byref(cam), C.byref(p), _p(prim), _p(t), _p(rad), _p(nr))
        return prim, t, rad, nr

    def sample_batch(self, points, p: Params):
        pts = np.ascontiguousarray(points, SAMPLE_POINT_DTYPE)
        out = np.empty(pts.shape[0], SAMPLED_COLOR_DTYPE)
        self._lib.orc_sample_batch(self._h, _p(pts), C.c_size_t(pts.shape[0]), C.byref(p), _p(out))
        return out

    def trace_rays(self, origins, dirs, t_min=0.001, use_bvh=True, want_normal=False):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3); d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        prim = np.empty(n, np.uint32); t = np.empty(n, np.float32); nrm = np.empty((n, 3), np.float32) if want_normal else None
        self._lib.orc_trace_rays(self._h, C.c_size_t(n), _p(o), _p(d), C.c_float(t_min), C.c_int(1 if use_bvh else 0), _p(prim), _p(t), _p(nrm))
        return (prim, t, nrm) if want_normal else (prim, t)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) for x in ctr]); k = (C.c_uint32 * 2)(*[int(x) for x in key]); o = (C.c_uint32 * 4)()
    load().orc_philox(c, k, o)
    return [int(x) for x in o]


def det_math(u):
    u = np.ascontiguousarray(u, np.float32)
    s, c, cb = np.empty_like(u), np.empty_like(u), np.empty_like(u)
    load().orc_det_math(C.c_size_t(u.shape[0]), _p(u), _p(s), _p(c), _p(cb))
    return s, c, cb


def make_ray(o, d):
    out = (C.c_float * 6)(); load().orc_make_ray(_f3(o), _f3(d), out); return np.array(out[:], np.float32)


def ray_at(ray6, t):
    r = (C.c_float * 6)(*[float(x) for x in ray6]); out = (C.c_float * 3)(); load().orc_ray_at(r, C.c_float(t), out)
    return np.array(out[:], np.float32)


def vec3_kat(a, b):
    ln, dt, cr = C.c_float(), C.c_float(), (C.c_float * 3)()
    load().orc_vec3_kat(_f3(a), _f3(b), C.byref(ln), C.byref(dt), cr)
    return ln.value, dt.value, np.array(cr[:], np.float32)


def scatter(kind, albedo, param, ray6, t, outward_normal, seed=1, pixel=0, sample=0, bounce=1, math_mode=MATH_DET):
    r = (C.c_float * 6)(*[float(x) for x in ray6]); o = (C.c_float * 6)(); a = (C.c_float * 3)()
    ok = load().orc_scatter(C.c_uint32(kind), _f3(albedo), C.c_float(param), r, C.c_float(t), _f3(outward_normal), C.c_uint64(seed),
                            C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(bounce), C.c_uint32(math_mode), o, a)
    return (bool(ok), np.array(o[:], np.float32), np.array(a[:], np.float32))


def tonemap(lin: float) -> int:
    return int(load().orc_tonemap(C.c_float(lin)))


def primary_ray(cam: Camera, x, y, seed, sample):
    out = (C.c_float * 6)()
    load().orc_primary_ray(C.byref(cam), C.c_uint32(x), C.c_uint32(y), C.c_uint64(seed), C.c_uint32(sample), out)
    return np.array(out[:], np.float32)
