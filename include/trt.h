/* include/trt.h — C ABI of libtrt_b200.so: the B200 (sm_100a) path-tracing backend behind tiny-raytracer's
 * scene / camera / material / render API.
 *
 * This is the drop-in boundary. Every entry point names the reference interface it replaces
 * (file:line relative to /root/reference/raytracer/src). Plain pointers and sizes only: no C++ or torch types.
 * All functions return TRT_OK (0) or a negative trt_status; trt_last_error() gives the message of the last failure
 * on the calling thread. Nothing unwinds across this boundary (the reference panics instead: sampler/cpu.rs:35,85).
 * Thread-compatibility: one trt_ctx per host thread, or lock externally. One trt_ctx drives one GPU. Multi-GPU: either
 * trt_multi_* / trt_render_multi below (one call, one process, threads + NVLink peer reduce inside the library), or one ctx
 * per process with the accumulation buffers reduced by the caller (NCCL; trt_render_accumulate + trt_finalize).
 * There is NO CPU fallback: every compute entry point fails with TRT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef TRT_B200_H
#define TRT_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TRT_ABI_VERSION 1

typedef enum trt_status {
    TRT_OK = 0,
    TRT_ERR_INVALID = -1,    /* bad argument / bad handle / bad state (e.g. render before commit) */
    TRT_ERR_CUDA = -2,       /* CUDA runtime error or no usable device */
    TRT_ERR_OOM = -3,
    TRT_ERR_UNSUPPORTED = -4
} trt_status;

typedef struct trt_ctx trt_ctx;       /* one GPU + one stream */
typedef struct trt_scene trt_scene;   /* replaces hittable::world::World (hittable/world.rs:10-13) */

/* How trt_scene_commit arranges the BVH (replaces BVH::new / Node::new, hittable/bvh.rs:12-21,42-84). */
typedef enum trt_bvh_kind {
    TRT_BVH_REFERENCE = 0,   /* the reference's median-split topology, traversed left-first: reference results by construction */
    TRT_BVH_SAH = 1,         /* binned-SAH topology, near-child-first traversal, ties broken by reference leaf order */
    TRT_BVH_FLAT = 2,        /* no tree: <= 64 primitives as a list in reference leaf order; reference results by construction */
    TRT_BVH_AUTO = 3,        /* FLAT when the scene has <= 64 primitives (fastest there, exact by construction), else SAH;
                                from 65 536 primitives on LBVH_DEVICE (same results, the build moves to the GPU) */
    TRT_BVH_LBVH_DEVICE = 4  /* built on the GPU during trt_scene_upload (Morton codes + radix sort + Karras), walked like SAH;
                                the reference's own Node::new also runs there (level-parallel), so ties and the replay of
                                order-dependent rays follow the reference leaf order exactly as with TRT_BVH_SAH */
} trt_bvh_kind;

/* Which kernel family renders. */
typedef enum trt_variant {
    TRT_VARIANT_MEGAKERNEL = 0,   /* persistent threads, per-lane path regeneration */
    TRT_VARIANT_WAVEFRONT = 1     /* ray queues in HBM, warp-ballot compaction between bounces */
} trt_variant;

/* Camera as plain data: the private fields of camera::Camera (camera.rs:4-14), produced by trt_camera_make. */
typedef struct trt_camera {
    float position[3], viewport_upper_left[3], forward[3], horizontal[3], vertical[3];
    float defocus_disk_u[3], defocus_disk_v[3];
    uint32_t width, height;
} trt_camera;

/* Replaces the arguments of Renderer::new (renderer/renderer.rs:21-35) plus the constants the reference hard-codes. */
typedef struct trt_params {
    uint32_t width, height;      /* must equal camera.width/height */
    uint32_t spp;                /* samples rendered by THIS call: global sample indices sample_offset .. +spp-1 */
    uint32_t sample_offset;
    uint32_t max_bounces;        /* Renderer::new max_bounces; the `while remain_bounces > 0` of sampler/cpu.rs:47 */
    float background[3];         /* Renderer::new background_color (None -> 0) */
    uint64_t seed;               /* Philox key; the reference's thread_rng() has no seed (utils/random.rs:15-18) */
    float t_min;                 /* 0.001 in the reference (sampler/cpu.rs:48) */
    uint32_t spp_total;          /* divisor of the final image; 0 -> spp. > spp when this call renders one rank's share */
    uint32_t variant;            /* trt_variant */
    uint32_t flags;              /* TRT_FLAG_* */
} trt_params;
#define TRT_FLAG_COUNT_EVENTS 1u   /* also count node visits / primitive tests (slower; for roofline accounting) */
#define TRT_FLAG_NO_TAIL_SPLIT 2u  /* megakernel: hand out whole pixels only (every pixel summed by one lane in sample order ->
                                    * run-to-run bit-identical f64 sums); default splits the last pixels into sample chunks
                                    * added with f64 atomics (tail balance; the f32 image can move by <= 1 ulp) */
#define TRT_FLAG_COUNT_WALKED 8u   /* count on the walk the scene really gets (4-wide records where it has them): node_visits = records
                                    * fetched (112 B each), sphere_tests / quad_tests = primitive records fetched. For the memory
                                    * roofline of big scenes; ignored when TRT_FLAG_COUNT_EVENTS is set. Results do not change. */
#define TRT_FLAG_INLINE_PRIMARY 16u /* megakernel: always generate and trace primary rays inside the path loop, never in the coherent
                                    * first-bounce pass (used by default only where it pays: wide-node trees in
                                    * shared memory). Same results bit for bit; A/B and tests. */
#define TRT_FLAG_PRIMARY_PASS 32u   /* megakernel: use the first-bounce pass whatever the scene (needs width*height*32 bytes of device
                                    * scratch per sample of a batch; loses to TRT_FLAG_INLINE_PRIMARY). Same results; A/B and tests. */
/* bits 16..17 of flags: step function of the step-wise megakernel loop on wide nodes in global memory (0 = chosen by the library,
 * 1 = 128-bit node loads, 2 = 256-bit node loads, 3 = 256-bit loads and one dependent fetch per step). Same results; A/B. */
#define TRT_FLAG_STEP(n) (((n) & 3u) << 16)
#define TRT_FLAG_NO_WIDE8 0x40000u  /* megakernel: walk the 4-wide nodes also where the scene has compressed 8-wide ones (built only when
                                    * TRT_WIDE8=1 is in the environment at trt_scene_upload, for ordered trees that do not fit shared
                                    * memory; DESIGN.md §4.5: measured slower on B200). Same results bit for bit; A/B and tests. */
#define TRT_FLAG_BINARY_NODES 4u   /* tree kinds: walk the binary nodes instead of the 4-wide ones (same results; A/B and tests) */
/* bits 8..15 of flags: megakernel tuning override, the number of lanes of a warp (1..32) that must have finished their ray
 * before the shade/regenerate block runs; 0 = chosen from the tree size (32 up to 4096 nodes, 16 above). Results do not
 * depend on it. */
#define TRT_FLAG_REGEN_MIN(n) (((uint32_t)(n) & 0xFFu) << 8)

typedef struct trt_stats {
    uint64_t paths, rays;                          /* pixel-samples; ray segments = iterations of sampler/cpu.rs:47 */
    uint64_t node_visits, sphere_tests, quad_tests;/* only with TRT_FLAG_COUNT_EVENTS */
    float device_ms;                               /* CUDA-event time of the render kernels of this call */
    float finalize_ms;                             /* normalise + tone-map kernel */
    uint32_t launches;                             /* kernels launched by this call */
    uint32_t reserved;
} trt_stats;

/* Wire structs of the reference's Sampler trait (renderer/pointgen.rs:7-13, renderer/imager.rs:9-15). */
typedef struct trt_sample_point { uint32_t x, y; float origin[3], direction[3]; } trt_sample_point;   /* 32 B */
typedef struct trt_sampled_color { uint32_t x, y; float color[3]; } trt_sampled_color;                /* 20 B */

const char* trt_last_error(void);
uint32_t trt_abi_version(void);

/* ---- context: one GPU. No reference counterpart (the reference has no device). */
int trt_ctx_create(int device_ordinal, trt_ctx** out);
int trt_ctx_destroy(trt_ctx* ctx);
int trt_ctx_device_name(trt_ctx* ctx, char* buf, size_t buf_len);
int trt_ctx_sm_count(trt_ctx* ctx, int* out);
int trt_ctx_synchronize(trt_ctx* ctx);
/* Run all later work of this ctx on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL = the legacy default
 * stream), e.g. the framework stream that also carries the NCCL reduce, so that ordering and CUDA-event timing line up. */
int trt_ctx_set_stream(trt_ctx* ctx, void* cuda_stream);

/* ---- scene building: World::new / add_material / add_geometry (hittable/world.rs:16-31). Host only, no GPU needed. */
int trt_scene_create(trt_scene** out);
int trt_scene_destroy(trt_scene* scene);
int trt_add_lambertian(trt_scene* scene, const float albedo[3], uint32_t* mat_id);          /* material/lambertian.rs:10-12 */
int trt_add_metal(trt_scene* scene, const float albedo[3], float fuzz, uint32_t* mat_id);   /* material/metal.rs:12-14 (clamps fuzz) */
int trt_add_dielectric(trt_scene* scene, const float albedo[3], float refraction_index, uint32_t* mat_id); /* material/dielectric.rs:12-14 */
int trt_add_light(trt_scene* scene, const float color[3], uint32_t* mat_id);                /* material/light.rs:10-12 */
int trt_add_sphere(trt_scene* scene, const float center[3], float radius, uint32_t mat_id, uint32_t* prim_id); /* hittable/sphere.rs:16-25 */
int trt_add_quad(trt_scene* scene, const float corner[3], const float u[3], const float v[3], uint32_t mat_id, uint32_t* prim_id); /* hittable/quad.rs:20-29 */
/* bulk add: type[i] 0 = sphere (data12: cx cy cz r), 1 = quad (data12: corner u v); 12 floats per primitive */
int trt_add_prims(trt_scene* scene, uint32_t n, const uint32_t* type, const float* data12, const uint32_t* mat_id);
int trt_scene_num_prims(trt_scene* scene, uint32_t* out);

/* World::get_bvh (hittable/world.rs:43-45): builds the tree on the host. Does not touch the GPU. */
int trt_scene_commit(trt_scene* scene, int bvh_kind);
/* left-to-right leaf order of the reference-topology tree, indexed by prim_id (after commit; for TRT_BVH_LBVH_DEVICE the
 * order is computed on the GPU and is available after trt_scene_upload, prim_id order before) */
int trt_scene_leaf_ranks(trt_scene* scene, uint32_t* rank_of_prim);
int trt_scene_bvh_info(trt_scene* scene, uint32_t* n_nodes, uint32_t* max_depth, float* sah_cost);
/* Seconds-scale costs of the last commit / upload: host build time of trt_scene_commit and, for TRT_BVH_LBVH_DEVICE, the
 * CUDA-event time of the build on the GPU (0 until uploaded). */
int trt_scene_build_times(trt_scene* scene, float* host_commit_ms, float* device_build_ms);
/* Copies the flattened scene into the GPU's HBM (replaces MetalAccelerationStructure::new, metal/accelstructure.rs:17-75). */
int trt_scene_upload(trt_ctx* ctx, trt_scene* scene);

/* Camera::new (camera.rs:17-56). Host only. */
int trt_camera_make(float focus_distance, float defocus_angle, const float position[3], const float look_at[3],
                    const float up[3], float vertical_fov, uint32_t width, uint32_t height, trt_camera* out);

/* ---- rendering: Renderer::render (renderer/renderer.rs:37-79) = generator + sampler + imager, all on the device.
 * HOST buffers in/out; copies are part of the call.
 *   out_linear [H*W*3] optional: (float)(sum64 * 1/spp_total)     (imager.rs:35,50)
 *   out_rgb8   [H*W*3] optional: gamma 2.2, clamp, *255 truncated  (utils/image.rs:92-111) */
int trt_render(trt_ctx* ctx, trt_scene* scene, const trt_camera* cam, const trt_params* p,
               float* out_linear, uint8_t* out_rgb8, trt_stats* stats);

/* Device-resident form used by multi-GPU callers and by the bench: accumulates the f64 per-pixel sums of this call's
 * samples into d_sum64 [H*W*3 doubles, DEVICE pointer, caller-zeroed]; no normalisation, no copies. */
int trt_render_accumulate(trt_ctx* ctx, trt_scene* scene, const trt_camera* cam, const trt_params* p,
                          double* d_sum64, trt_stats* stats);
/* Epilogue (imager.rs:35-60 + utils/image.rs:92-111) on DEVICE pointers: d_linear/d_rgb8 optional. */
int trt_finalize(trt_ctx* ctx, const double* d_sum64, uint32_t width, uint32_t height, uint32_t spp_total,
                 float* d_linear, uint8_t* d_rgb8, trt_stats* stats);

/* ---- the same render on several GPUs of one box (Renderer::render, renderer/renderer.rs:37-79, is still the one
 * call the user makes). One host thread + one trt_ctx per listed device (a device may be listed more than once); the samples
 * [p->sample_offset, +p->spp) are partitioned contiguously across them (Philox is keyed by the global sample index, so the set
 * of paths is the 1-GPU set); every GPU accumulates f64 sums of its share; then ONE fused kernel per GPU reduce-scatters over
 * NVLink peer memory — GPU k reads slice k of every peer's buffer, adds in rank order — and normalises + tone-maps its slice,
 * which is copied straight into the caller's HOST buffers (out_linear / out_rgb8 [H*W*3], either may be NULL; both NULL = leave
 * the result on the devices, for timing). Equals trt_render to <= 1 ulp (f64 partial sums in another order). No NCCL involved.
 * stats: paths / rays / events summed over the devices; device_ms and finalize_ms are the MAX over the devices (they run side
 * by side). Without peer access between two listed devices the slices are staged with cudaMemcpyPeer instead (same result). */
typedef struct trt_multi trt_multi;
int trt_multi_create(const int* devices, int n_devices, trt_multi** out);      /* 1..16 devices */
int trt_multi_destroy(trt_multi* multi);
int trt_multi_device_count(trt_multi* multi, int* n_devices, int* peer_direct /* optional: 1 = peer loads, 0 = staged copies */);
int trt_multi_render(trt_multi* multi, trt_scene* scene, const trt_camera* cam, const trt_params* p,
                     float* out_linear, uint8_t* out_rgb8, trt_stats* stats);
/* create + render + destroy in one call */
int trt_render_multi(const int* devices, int n_devices, trt_scene* scene, const trt_camera* cam, const trt_params* p,
                     float* out_linear, uint8_t* out_rgb8, trt_stats* stats);

/* Per-sample AOVs, HOST buffers indexed [pixel*spp + i]: primary-hit prim_id (0xFFFFFFFF = miss) and t, optional
 * whole-path radiance[3] and ray count. No reference counterpart (the reference has no primitive id); parity hook. */
int trt_render_aov(trt_ctx* ctx, trt_scene* scene, const trt_camera* cam, const trt_params* p,
                   uint32_t* prim_id, float* t, float* radiance, uint32_t* n_rays);

/* Sampler::sampling (renderer/sampler/mod.rs:9-16) on a batch: SamplePoint in, SampledColor out, same order. HOST buffers.
 * Item i uses the RNG stream (pixel = y*width + x, sample = sample_offset + i). */
int trt_sample_batch(trt_ctx* ctx, trt_scene* scene, const trt_sample_point* in, size_t n, const trt_params* p,
                     trt_sampled_color* out);

/* Streaming form of the same adapter: what MetalSampler::add_command / sampling do with command
 * buffers (renderer/sampler/metal/sampler.rs:43-88,107-130). Two slots of pinned host + device buffers on two streams:
 * submit() queues H2D copy + kernel + D2H copy of one batch (<= batch_capacity samples; the reference's batch is 10 240,
 * metal/sampler.rs:11) and returns; collect() waits for the OLDEST batch in flight and returns its colours in submission
 * order. At most two batches are in flight: a third submit() fails with TRT_ERR_INVALID until one is collected.
 * Item i of the whole stream uses the RNG stream (pixel, sample = params.sample_offset + i), so streaming equals one
 * trt_sample_batch call over the concatenated input. */
typedef struct trt_sampler trt_sampler;
int trt_sampler_create(trt_ctx* ctx, trt_scene* scene, const trt_params* p, size_t batch_capacity, trt_sampler** out);
int trt_sampler_submit(trt_sampler* sampler, const trt_sample_point* in, size_t n);
int trt_sampler_collect(trt_sampler* sampler, trt_sampled_color* out, size_t* n_out);
int trt_sampler_destroy(trt_sampler* sampler);

/* Closest hit for arbitrary rays (Hittable::hit on the BVH, hittable/bvh.rs:25-27), HOST buffers, directions as given. */
int trt_trace_rays(trt_ctx* ctx, trt_scene* scene, size_t n, const float* origins, const float* dirs, float t_min,
                   uint32_t* prim_id, float* t);

/* Device evaluation of the numeric contract for unit parity tests: HOST buffers of n floats. */
int trt_debug_math(trt_ctx* ctx, size_t n, const float* u, float* sin2pi, float* cos2pi, float* cbrt_out);
/* Philox4x32-10 on the device: n counters (4 words each) with one key (2 words) -> n*4 words. */
int trt_debug_philox(trt_ctx* ctx, size_t n, const uint32_t* ctr, const uint32_t key[2], uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* TRT_B200_H */
