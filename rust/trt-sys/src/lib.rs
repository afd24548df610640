//! Raw bindings of include/trt.h. SOURCE ONLY (no Rust toolchain in this image); kept in lock-step with the header by
//! tests/test_capi_host.py::test_rust_bindings_name_every_symbol.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct trt_ctx { _p: [u8; 0] }
#[repr(C)] pub struct trt_scene { _p: [u8; 0] }
#[repr(C)] pub struct trt_sampler { _p: [u8; 0] }
#[repr(C)] pub struct trt_multi { _p: [u8; 0] }

pub const TRT_OK: c_int = 0;
pub const TRT_BVH_REFERENCE: c_int = 0;
pub const TRT_BVH_SAH: c_int = 1;
pub const TRT_BVH_FLAT: c_int = 2;
pub const TRT_BVH_AUTO: c_int = 3;
pub const TRT_BVH_LBVH_DEVICE: c_int = 4;
pub const TRT_VARIANT_MEGAKERNEL: u32 = 0;
pub const TRT_VARIANT_WAVEFRONT: u32 = 1;
pub const TRT_FLAG_COUNT_EVENTS: u32 = 1;
pub const TRT_FLAG_NO_TAIL_SPLIT: u32 = 2;
pub const TRT_FLAG_BINARY_NODES: u32 = 4;
pub const TRT_FLAG_COUNT_WALKED: u32 = 8;
pub const TRT_FLAG_INLINE_PRIMARY: u32 = 16;
pub const TRT_FLAG_PRIMARY_PASS: u32 = 32;
pub const TRT_FLAG_NO_WIDE8: u32 = 0x40000;
pub const fn trt_flag_step(n: u32) -> u32 { (n & 3) << 16 }
pub const fn trt_flag_regen_min(n: u32) -> u32 { (n & 0xFF) << 8 }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct trt_camera {
    pub position: [f32; 3], pub viewport_upper_left: [f32; 3], pub forward: [f32; 3], pub horizontal: [f32; 3],
    pub vertical: [f32; 3], pub defocus_disk_u: [f32; 3], pub defocus_disk_v: [f32; 3], pub width: u32, pub height: u32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct trt_params {
    pub width: u32, pub height: u32, pub spp: u32, pub sample_offset: u32, pub max_bounces: u32, pub background: [f32; 3],
    pub seed: u64, pub t_min: f32, pub spp_total: u32, pub variant: u32, pub flags: u32,
}
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct trt_stats {
    pub paths: u64, pub rays: u64, pub node_visits: u64, pub sphere_tests: u64, pub quad_tests: u64,
    pub device_ms: f32, pub finalize_ms: f32, pub launches: u32, pub reserved: u32,
}
/// `#[repr(C)] SamplePoint` of raytracer/src/renderer/pointgen.rs:7-13 (32 B)
#[repr(C)] #[derive(Clone, Copy)] pub struct trt_sample_point { pub x: u32, pub y: u32, pub origin: [f32; 3], pub direction: [f32; 3] }
/// `#[repr(C)] SampledColor` of raytracer/src/renderer/imager.rs:9-15 (20 B)
#[repr(C)] #[derive(Clone, Copy)] pub struct trt_sampled_color { pub x: u32, pub y: u32, pub color: [f32; 3] }

extern "C" {
    pub fn trt_last_error() -> *const c_char;
    pub fn trt_abi_version() -> u32;
    pub fn trt_ctx_create(device_ordinal: c_int, out: *mut *mut trt_ctx) -> c_int;
    pub fn trt_ctx_destroy(ctx: *mut trt_ctx) -> c_int;
    pub fn trt_ctx_device_name(ctx: *mut trt_ctx, buf: *mut c_char, buf_len: usize) -> c_int;
    pub fn trt_ctx_sm_count(ctx: *mut trt_ctx, out: *mut c_int) -> c_int;
    pub fn trt_ctx_synchronize(ctx: *mut trt_ctx) -> c_int;
    pub fn trt_ctx_set_stream(ctx: *mut trt_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn trt_scene_create(out: *mut *mut trt_scene) -> c_int;
    pub fn trt_scene_destroy(scene: *mut trt_scene) -> c_int;
    pub fn trt_add_lambertian(scene: *mut trt_scene, albedo: *const f32, mat_id: *mut u32) -> c_int;
    pub fn trt_add_metal(scene: *mut trt_scene, albedo: *const f32, fuzz: f32, mat_id: *mut u32) -> c_int;
    pub fn trt_add_dielectric(scene: *mut trt_scene, albedo: *const f32, refraction_index: f32, mat_id: *mut u32) -> c_int;
    pub fn trt_add_light(scene: *mut trt_scene, color: *const f32, mat_id: *mut u32) -> c_int;
    pub fn trt_add_sphere(scene: *mut trt_scene, center: *const f32, radius: f32, mat_id: u32, prim_id: *mut u32) -> c_int;
    pub fn trt_add_quad(scene: *mut trt_scene, corner: *const f32, u: *const f32, v: *const f32, mat_id: u32, prim_id: *mut u32) -> c_int;
    pub fn trt_add_prims(scene: *mut trt_scene, n: u32, type_: *const u32, data12: *const f32, mat_id: *const u32) -> c_int;
    pub fn trt_scene_num_prims(scene: *mut trt_scene, out: *mut u32) -> c_int;
    pub fn trt_scene_commit(scene: *mut trt_scene, bvh_kind: c_int) -> c_int;
    pub fn trt_scene_leaf_ranks(scene: *mut trt_scene, rank_of_prim: *mut u32) -> c_int;
    pub fn trt_scene_bvh_info(scene: *mut trt_scene, n_nodes: *mut u32, max_depth: *mut u32, sah_cost: *mut f32) -> c_int;
    pub fn trt_scene_build_times(scene: *mut trt_scene, host_commit_ms: *mut f32, device_build_ms: *mut f32) -> c_int;
    pub fn trt_scene_upload(ctx: *mut trt_ctx, scene: *mut trt_scene) -> c_int;
    pub fn trt_camera_make(focus_distance: f32, defocus_angle: f32, position: *const f32, look_at: *const f32, up: *const f32,
                           vertical_fov: f32, width: u32, height: u32, out: *mut trt_camera) -> c_int;
    pub fn trt_render(ctx: *mut trt_ctx, scene: *mut trt_scene, cam: *const trt_camera, p: *const trt_params,
                      out_linear: *mut f32, out_rgb8: *mut u8, stats: *mut trt_stats) -> c_int;
    pub fn trt_render_accumulate(ctx: *mut trt_ctx, scene: *mut trt_scene, cam: *const trt_camera, p: *const trt_params,
                                 d_sum64: *mut f64, stats: *mut trt_stats) -> c_int;
    pub fn trt_finalize(ctx: *mut trt_ctx, d_sum64: *const f64, width: u32, height: u32, spp_total: u32,
                        d_linear: *mut f32, d_rgb8: *mut u8, stats: *mut trt_stats) -> c_int;
    pub fn trt_multi_create(devices: *const c_int, n_devices: c_int, out: *mut *mut trt_multi) -> c_int;
    pub fn trt_multi_destroy(multi: *mut trt_multi) -> c_int;
    pub fn trt_multi_device_count(multi: *mut trt_multi, n_devices: *mut c_int, peer_direct: *mut c_int) -> c_int;
    pub fn trt_multi_render(multi: *mut trt_multi, scene: *mut trt_scene, cam: *const trt_camera, p: *const trt_params,
                            out_linear: *mut f32, out_rgb8: *mut u8, stats: *mut trt_stats) -> c_int;
    pub fn trt_render_multi(devices: *const c_int, n_devices: c_int, scene: *mut trt_scene, cam: *const trt_camera, p: *const trt_params,
                            out_linear: *mut f32, out_rgb8: *mut u8, stats: *mut trt_stats) -> c_int;
    pub fn trt_render_aov(ctx: *mut trt_ctx, scene: *mut trt_scene, cam: *const trt_camera, p: *const trt_params,
                          prim_id: *mut u32, t: *mut f32, radiance: *mut f32, n_rays: *mut u32) -> c_int;
    pub fn trt_sample_batch(ctx: *mut trt_ctx, scene: *mut trt_scene, input: *const trt_sample_point, n: usize,
                            p: *const trt_params, out: *mut trt_sampled_color) -> c_int;
    pub fn trt_sampler_create(ctx: *mut trt_ctx, scene: *mut trt_scene, p: *const trt_params, batch_capacity: usize, out: *mut *mut trt_sampler) -> c_int;
    pub fn trt_sampler_submit(sampler: *mut trt_sampler, input: *const trt_sample_point, n: usize) -> c_int;
    pub fn trt_sampler_collect(sampler: *mut trt_sampler, out: *mut trt_sampled_color, n_out: *mut usize) -> c_int;
    pub fn trt_sampler_destroy(sampler: *mut trt_sampler) -> c_int;
    pub fn trt_trace_rays(ctx: *mut trt_ctx, scene: *mut trt_scene, n: usize, origins: *const f32, dirs: *const f32, t_min: f32,
                          prim_id: *mut u32, t: *mut f32) -> c_int;
    pub fn trt_debug_math(ctx: *mut trt_ctx, n: usize, u: *const f32, sin2pi: *mut f32, cos2pi: *mut f32, cbrt_out: *mut f32) -> c_int;
    pub fn trt_debug_philox(ctx: *mut trt_ctx, n: usize, ctr: *const u32, key: *const u32, out: *mut u32) -> c_int;
}
