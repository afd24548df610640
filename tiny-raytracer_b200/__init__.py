"""tiny-raytracer_b200 — B200-native path-tracing backend behind tiny-raytracer's scene/camera/material/render API.

The directory name follows the reference repo's name; Python imports it as `tiny_raytracer_b200` through the shim
module of that name at the repository root. All pixels are computed by csrc/ (CUDA, sm_100a) via libtrt_b200.so.
"""
from . import _lib
from ._lib import TrtError, LIB_PATH
from .api import (SamplerStream, Context, MultiContext, Scene, make_camera, make_params, finalize, debug_math, debug_philox,
                  Vec3, Transform, Lambertian, Metal, Dielectric, Light, Sphere, Quad, World, Camera, Renderer, Image, new_box,
                  BVH_REFERENCE, BVH_SAH, BVH_FLAT, BVH_AUTO, BVH_LBVH_DEVICE, VARIANT_MEGAKERNEL, VARIANT_WAVEFRONT, SAMPLE_POINT_DTYPE, SAMPLED_COLOR_DTYPE)
from . import scenes
