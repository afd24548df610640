//! The reference-side binding a maintainer would add: `CudaRenderer` has the signature of
//! `raytracer::renderer::Renderer` (renderer/renderer.rs:21-35,37) and `CudaSampler` implements the `Sampler` trait
//! (renderer/sampler/mod.rs:9-16) by batching, like the reference's own MetalSampler (metal/sampler.rs:91-132).
//! SOURCE ONLY: never compiled here (no Rust toolchain; the reference crate itself is macOS-only).
//! It needs two small accessors in the reference, because its material and camera fields are private
//! (lambertian.rs:5-7, metal.rs:6-9, dielectric.rs:6-9, light.rs:6-8, camera.rs:4-14):
//!     impl Material { fn to_pod(&self) -> (u32 /*kind*/, Vec3 /*albedo|colour*/, f32 /*fuzz|ior*/) }
//!     impl Camera   { fn args(&self) -> (f32, f32, Vec3, Vec3, Vec3, f32, usize, usize) }   // what Camera::new was given
//!     impl World    { pub fn geometries_in_order(&self) -> &[Arc<Box<dyn Hittable>>] }      // insertion order across types
use std::{ffi::CStr, ptr, sync::Arc};
use as_any::Downcast;
use raytracer::{camera::Camera, hittable::{quad::Quad, sphere::Sphere, world::World}, math::vec3::Vec3,
                renderer::{imager::SampledColor, pointgen::SamplePoint, sampler::Sampler}, utils::image::{Color, Image}};
use tokio::task::JoinHandle;
use trt_sys::*;

fn check(rc: i32) { if rc != TRT_OK { panic!("{}", unsafe { CStr::from_ptr(trt_last_error()) }.to_string_lossy()); } }   // the reference panics too
fn v(a: Vec3) -> [f32; 3] { [a.x, a.y, a.z] }

/// World -> trt_scene. Primitive ids — and, through the stable sort of `Node::new` (hittable/bvh.rs:62-66), the reference leaf
/// order that decides exact-t ties — must follow the order in which `World::add_geometry` was called ACROSS types. The per-type
/// hook `get_geometries::<T>()` (hittable/world.rs:47-55) loses that order for mixed scenes, so this binding uses the third
/// accessor INTEGRATION.md §2 proposes for the reference (a read-only view of the private `geometries` vector):
///     impl World { pub fn geometries_in_order(&self) -> &[Arc<Box<dyn Hittable>>] { &self.geometries } }
/// and downcasts every entry itself, exactly as `get_geometries` does (`(***object).downcast_ref::<T>()`, world.rs:50).
unsafe fn upload_world(world: &World, scene: *mut trt_scene) {
    let mut ids = std::collections::HashMap::new();
    let mut mat = |m: &Arc<Box<dyn raytracer::material::Material>>| -> u32 {
        *ids.entry(Arc::as_ptr(m) as *const u8 as usize).or_insert_with(|| {
            let (kind, rgb, param) = m.to_pod();
            let mut id = 0u32;
            check(match kind {
                0 => trt_add_lambertian(scene, v(rgb).as_ptr(), &mut id),
                1 => trt_add_metal(scene, v(rgb).as_ptr(), param, &mut id),
                2 => trt_add_dielectric(scene, v(rgb).as_ptr(), param, &mut id),
                _ => trt_add_light(scene, v(rgb).as_ptr(), &mut id),
            });
            id
        })
    };
    for object in world.geometries_in_order() {
        if let Some(q) = (***object).downcast_ref::<Quad>() {
            let m = mat(&q.material);
            check(trt_add_quad(scene, v(q.corner).as_ptr(), v(q.u).as_ptr(), v(q.v).as_ptr(), m, ptr::null_mut()));
        } else if let Some(s) = (***object).downcast_ref::<Sphere>() {
            let m = mat(&s.material);
            check(trt_add_sphere(scene, v(s.center).as_ptr(), s.radius, m, ptr::null_mut()));
        } else {
            panic!("raytracer-cuda: unsupported Hittable (only Sphere and Quad exist in the reference, hittable/mod.rs:7-11)");
        }
    }
}

#[derive(Clone, Copy)]
pub struct CudaRenderer { samples_per_pixel: usize, max_bounces: usize, background_color: Vec3, pub seed: u64, pub device: i32,
                          /// GPUs of the box to render on (devices 0..n_gpus): > 1 goes through trt_render_multi — samples-per-pixel
                          /// partitioned, NVLink peer reduce + tone map on the GPUs, one call
                          pub n_gpus: i32 }

impl CudaRenderer {
    /// Same arguments as `Renderer::new`; `num_sampler_threads` and `progressbar` are accepted and ignored.
    pub fn new(samples_per_pixel: usize, _num_sampler_threads: usize, max_bounces: usize, _progressbar: bool,
               background_color: Option<Vec3>) -> Self {
        Self { samples_per_pixel, max_bounces, background_color: background_color.unwrap_or(Vec3::zero()), seed: 1, device: 0, n_gpus: 1 }
    }

    /// Same signature as `Renderer::render`: generator + sampler + imager all run on the GPU inside `trt_render`.
    pub fn render(&self, camera: Camera, world: Arc<World>) -> JoinHandle<Image> {
        let this = *self;
        tokio::task::spawn_blocking(move || unsafe {
            let (mut ctx, mut scene) = (ptr::null_mut(), ptr::null_mut());
            if this.n_gpus <= 1 { check(trt_ctx_create(this.device, &mut ctx)); }
            check(trt_scene_create(&mut scene));
            upload_world(&world, scene);
            check(trt_scene_commit(scene, TRT_BVH_AUTO));
            let (fd, da, pos, at, up, fov, w, h) = camera.args();
            let mut cam = trt_camera::default();
            check(trt_camera_make(fd, da, v(pos).as_ptr(), v(at).as_ptr(), v(up).as_ptr(), fov, w as u32, h as u32, &mut cam));
            let p = trt_params { width: w as u32, height: h as u32, spp: this.samples_per_pixel as u32, sample_offset: 0,
                                 max_bounces: this.max_bounces as u32, background: v(this.background_color), seed: this.seed,
                                 t_min: 0.001, spp_total: this.samples_per_pixel as u32, variant: TRT_VARIANT_MEGAKERNEL, flags: 0 };
            let mut linear = vec![0f32; w * h * 3];
            if this.n_gpus > 1 {
                let devices: Vec<i32> = (0..this.n_gpus).collect();
                check(trt_render_multi(devices.as_ptr(), this.n_gpus, scene, &cam, &p, linear.as_mut_ptr(), ptr::null_mut(), ptr::null_mut()));
            } else {
                check(trt_render(ctx, scene, &cam, &p, linear.as_mut_ptr(), ptr::null_mut(), ptr::null_mut()));
            }
            trt_scene_destroy(scene); trt_ctx_destroy(ctx);
            let mut image = Image::new_with_gamma_correction(w, h, 2.2);          // imager.rs:36-40
            for y in 0..h { for x in 0..w { let i = (y * w + x) * 3; image.set_pixel(x, y, Color::new(linear[i], linear[i + 1], linear[i + 2])); } }
            image
        })
    }
}

/// Trait-level adapter: batches of 10 240 SamplePoints (the reference's own GPU batch size, metal/sampler.rs:11).
pub struct CudaSampler { pub max_bounces: usize, pub background_color: Vec3, pub seed: u64, pub width: u32, pub height: u32 }

impl Sampler for CudaSampler {
    async fn sampling(self, world: &World, in_channel: flume::Receiver<SamplePoint>, out_channel: flume::Sender<SampledColor>) {
        unsafe {
            let (mut ctx, mut scene) = (ptr::null_mut(), ptr::null_mut());
            check(trt_ctx_create(0, &mut ctx)); check(trt_scene_create(&mut scene));
            upload_world(world, scene); check(trt_scene_commit(scene, TRT_BVH_AUTO));
            let mut p = trt_params { width: self.width, height: self.height, spp: 1, sample_offset: 0, max_bounces: self.max_bounces as u32,
                                     background: v(self.background_color), seed: self.seed, t_min: 0.001, spp_total: 1,
                                     variant: TRT_VARIANT_MEGAKERNEL, flags: 0 };
            let mut batch: Vec<SamplePoint> = Vec::with_capacity(10240);
            loop {
                let next = in_channel.recv_async().await;
                if let Ok(sp) = next { batch.push(sp); }
                if batch.len() == 10240 || (next.is_err() && !batch.is_empty()) {
                    let mut out = vec![std::mem::zeroed::<SampledColor>(); batch.len()];
                    // SamplePoint / SampledColor are #[repr(C)] and layout-identical to trt_sample_point / trt_sampled_color
                    check(trt_sample_batch(ctx, scene, batch.as_ptr() as *const trt_sample_point, batch.len(), &p,
                                           out.as_mut_ptr() as *mut trt_sampled_color));
                    p.sample_offset += batch.len() as u32;
                    for c in out { out_channel.send_async(c).await.expect("failed to send sampled color"); }
                    batch.clear();
                }
                if next.is_err() { break; }
            }
            trt_scene_destroy(scene); trt_ctx_destroy(ctx);
        }   // out_channel dropped here so that Imager::collect terminates (imager.rs:46)
    }
}
