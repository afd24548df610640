"""ctypes binding of libtrt_b200.so (include/trt.h). Loading fails loudly: there is no Python or CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRT_B200_LIB") or os.path.join(_HERE, "libtrt_b200.so")   # override: A/B builds of the same ABI

TRT_OK = 0
BVH_REFERENCE, BVH_SAH, BVH_FLAT, BVH_AUTO, BVH_LBVH_DEVICE = 0, 1, 2, 3, 4
VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT = 0, 1
FLAG_COUNT_EVENTS = 1
FLAG_NO_TAIL_SPLIT = 2
FLAG_BINARY_NODES = 4
FLAG_COUNT_WALKED = 8
FLAG_INLINE_PRIMARY = 16
FLAG_PRIMARY_PASS = 32
FLAG_NO_WIDE8 = 0x40000


def flag_step(n):
    """Step function of the step-wise loop on wide nodes in global memory (trt.h TRT_FLAG_STEP)."""
    return (n & 3) << 16


class TrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("trt error %d: %s" % (code, msg))
        self.code = code


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("viewport_upper_left", C.c_float * 3), ("forward", C.c_float * 3),
                ("horizontal", C.c_float * 3), ("vertical", C.c_float * 3), ("defocus_disk_u", C.c_float * 3),
                ("defocus_disk_v", C.c_float * 3), ("width", C.c_uint32), ("height", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("spp", C.c_uint32), ("sample_offset", C.c_uint32),
                ("max_bounces", C.c_uint32), ("background", C.c_float * 3), ("seed", C.c_uint64), ("t_min", C.c_float),
                ("spp_total", C.c_uint32), ("variant", C.c_uint32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("node_visits", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("quad_tests", C.c_uint64), ("device_ms", C.c_float), ("finalize_ms", C.c_float), ("launches", C.c_uint32),
                ("reserved", C.c_uint32)]


class SamplePoint(C.Structure):   # renderer/pointgen.rs:7-13, 32 B
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("origin", C.c_float * 3), ("direction", C.c_float * 3)]


class SampledColor(C.Structure):  # renderer/imager.rs:9-15, 20 B
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("color", C.c_float * 3)]


# every symbol include/trt.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F3 = C.POINTER(C.c_float)
_U32P = C.POINTER(C.c_uint32)
SYMBOLS = {
    "trt_last_error": (C.c_char_p, []),
    "trt_abi_version": (C.c_uint32, []),
    "trt_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "trt_ctx_destroy": (C.c_int, [_P]),
    "trt_ctx_device_name": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
    "trt_ctx_sm_count": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "trt_ctx_synchronize": (C.c_int, [_P]),
    "trt_ctx_set_stream": (C.c_int, [_P, _P]),
    "trt_scene_create": (C.c_int, [C.POINTER(_P)]),
    "trt_scene_destroy": (C.c_int, [_P]),
    "trt_add_lambertian": (C.c_int, [_P, _F3, _U32P]),
    "trt_add_metal": (C.c_int, [_P, _F3, C.c_float, _U32P]),
    "trt_add_dielectric": (C.c_int, [_P, _F3, C.c_float, _U32P]),
    "trt_add_light": (C.c_int, [_P, _F3, _U32P]),
    "trt_add_sphere": (C.c_int, [_P, _F3, C.c_float, C.c_uint32, _U32P]),
    "trt_add_quad": (C.c_int, [_P, _F3, _F3, _F3, C.c_uint32, _U32P]),
    "trt_add_prims": (C.c_int, [_P, C.c_uint32, _P, _P, _P]),
    "trt_scene_num_prims": (C.c_int, [_P, _U32P]),
    "trt_scene_commit": (C.c_int, [_P, C.c_int]),
    "trt_scene_leaf_ranks": (C.c_int, [_P, _P]),
    "trt_scene_bvh_info": (C.c_int, [_P, _U32P, _U32P, _F3]),
    "trt_scene_build_times": (C.c_int, [_P, _F3, _F3]),
    "trt_scene_upload": (C.c_int, [_P, _P]),
    "trt_camera_make": (C.c_int, [C.c_float, C.c_float, _F3, _F3, _F3, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(Camera)]),
    "trt_render": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Params), _P, _P, C.POINTER(Stats)]),
    "trt_render_accumulate": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Params), _P, C.POINTER(Stats)]),
    "trt_finalize": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, C.POINTER(Stats)]),
    "trt_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "trt_multi_destroy": (C.c_int, [_P]),
    "trt_multi_device_count": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "trt_multi_render": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Params), _P, _P, C.POINTER(Stats)]),
    "trt_render_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, _P, C.POINTER(Camera), C.POINTER(Params), _P, _P, C.POINTER(Stats)]),
    "trt_render_aov": (C.c_int, [_P, _P, C.POINTER(Camera), C.POINTER(Params), _P, _P, _P, _P]),
    "trt_sample_batch": (C.c_int, [_P, _P, _P, C.c_size_t, C.POINTER(Params), _P]),
    "trt_sampler_create": (C.c_int, [_P, _P, C.POINTER(Params), C.c_size_t, C.POINTER(_P)]),
    "trt_sampler_submit": (C.c_int, [_P, _P, C.c_size_t]),
    "trt_sampler_collect": (C.c_int, [_P, _P, C.POINTER(C.c_size_t)]),
    "trt_sampler_destroy": (C.c_int, [_P]),
    "trt_trace_rays": (C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_float, _P, _P]),
    "trt_debug_math": (C.c_int, [_P, C.c_size_t, _P, _P, _P, _P]),
    "trt_debug_philox": (C.c_int, [_P, C.c_size_t, _P, _U32P, _P]),
}

_lib = None


def load():
    """Load libtrt_b200.so, binding every declared symbol. Raises if the library or any symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(tiny-raytracer_b200/csrc/build.sh). There is no fallback implementation." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != TRT_OK:
        raise TrtError(rc, load().trt_last_error().decode("utf-8", "replace"))


def f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))
