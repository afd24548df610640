// trt_kernels.cu — sm_100a kernels of the path-tracing hot path and their launchers.
//
//   k_render_mega / k_render_mega_flat
//                   persistent-thread megakernel: every lane owns a work item (pixel, sample chunk) and alternates
//                   T (one closest-hit query) and S (shade-or-regenerate); the chunk's f64 sum stays in registers and
//                   is added to the per-pixel accumulator once. The _flat entry carries the primitive list of a
//                   TRT_BVH_FLAT scene in its kernel parameters (constant bank) for the uniform phase-A loop.
//   k_aov           one thread per (pixel, sample): primary-hit AOV (+ optional whole-path radiance / ray count)
//   k_sample_batch  the reference's Sampler trait on a batch of caller-supplied rays
//   k_trace         closest hit for arbitrary rays
//   k_finalize      normalise + gamma 2.2 + 8-bit quantise (imager.rs:35-60, utils/image.rs:92-111)
//   k_debug_*       device evaluation of the numeric contract for unit parity tests
#include "trt_device.cuh"
#include "trt_kernels.h"
#include "trt_host_glue.cuh"
#include <cstdio>

using namespace trtd;

#ifndef TRT_MEGA_THREADS
#define TRT_MEGA_THREADS 256
#endif
#ifndef TRT_MEGA_MIN_BLOCKS
#define TRT_MEGA_MIN_BLOCKS 4      // 64 registers, ~120 B of spills that stay in L1: measured +6 % (Cornell) and +14 % (1 M spheres) over 3 blocks at 80 registers
#endif
#ifndef TRT_FLAT_MIN_BLOCKS
#define TRT_FLAT_MIN_BLOCKS 3   // 80 registers, no spills: measured equal to 4 blocks with spills, and ahead of 5
#endif
#define MEGA_THREADS TRT_MEGA_THREADS
#ifndef TRT_FLAT_THREADS
#define TRT_FLAT_THREADS 256
#endif

// ------------------------------------------------------------------ megakernel
struct MegaArgs {
    DevScene scene; DevCamera cam; DevParams p;
    double* sum64;                 // [H*W*3]
    unsigned long long* counters;  // [0] rays [1] paths [2] node_visits [3] sphere_tests [4] quad_tests
    unsigned int* next_item;       // work counter
    uint32_t chunk, n_chunks, tiles_x, tiles_y, n_items;
    uint32_t tiles_whole;          // the first tiles_whole tiles are handed out as whole pixels (all samples, plain store);
                                   // the remaining tiles as n_chunks items of `chunk` samples each (f64 atomicAdd): small
                                   // items at the end of the job, so that lanes do not idle behind the last whole pixel
    uint32_t regen_min;            // lanes that must be waiting before the shade/regenerate block runs (32 = T-then-S)
    const float4* prim_a; const float4* prim_b;   // PRIM kernels: the records of k_primary for this launch's samples, [spp][H*W] x 2 float4 (prim_b unused)
};
#ifndef TRT_STEP_DEFAULT
#define TRT_STEP_DEFAULT 2      /* measured on 10^6 spheres (B200): 1 -> 2 117, 2 -> 2 328, 3 -> 2 303 Mrays/s (device-built tree); 2 242 / 2 458 / 2 430 (SAH) */
#endif
struct FlatList { float4 cbox[TRT_FLAT_CBOX_F4]; };   // phase-A boxes of the flat list, read as kernel parameters (constant bank)

// item -> (pixel, sample range). 32 consecutive items = one 8x4 pixel tile. `exclusive` = this lane alone adds to the pixel.
__device__ __forceinline__ bool decode_item(const MegaArgs& A, uint32_t item, uint32_t* x, uint32_t* y, uint32_t* s0, uint32_t* s1, bool* exclusive) {
    uint32_t lane = item & 31u, t = item >> 5, tile;
    if (t < A.tiles_whole) { tile = t; *s0 = 0; *s1 = A.p.spp; *exclusive = true; }
    else {
        uint32_t r = t - A.tiles_whole, tiles_rest = A.tiles_x * A.tiles_y - A.tiles_whole;
        uint32_t chunk = r / tiles_rest;                      // chunk-major over the remaining tiles
        tile = A.tiles_whole + (r - chunk * tiles_rest);
        *s0 = chunk * A.chunk; uint32_t e = *s0 + A.chunk; *s1 = e < A.p.spp ? e : A.p.spp;
        *exclusive = A.n_chunks == 1;
    }
    uint32_t ty = tile / A.tiles_x, tx = tile - ty * A.tiles_x;
    *x = tx * 8 + (lane & 7u); *y = ty * 4 + (lane >> 3);
    return *x < A.p.width && *y < A.p.height;
}

// ------------------------------------------------------------------ the coherent first-bounce pass (round 2)
// Primary rays are the one coherent population of a path tracer: the 32 lanes of a warp take the 32 pixels of an 8x4 tile at
// the SAME sample index, so camera-ray generation runs with all lanes (inside the path loop it ran for the ~7 lanes that
// happened to start a path in an iteration: 144 instructions at 5 lanes) and the closest-hit query sees nearly identical rays
// (same candidates / same subtree for the whole warp). k_primary does pointgen.rs:37-52 + camera.rs:58-66 + the first
// BVH::hit of cpu.rs:48 for a batch of samples of every pixel and leaves one 32-byte record per path in HBM:
//     a = (origin.xyz, t)    b = (direction.xyz, hit ref bits)
// A path that is already over after its first segment — a miss (background), a light (emission), or max_bounces <= 1 — is
// stored as DONE with its final colour in a.xyz (computed with the path loop's own operations: 0 + 1 * c), so the path loop
// only adds it up. The path loop (mega_body*, PRIM = true) starts every path from its record: no camera code, no first
// traversal, and a whole-background warp consumes records at a few instructions per path. Same arithmetic, same Philox
// streams, same summation order: results are bit-identical to the inline form (TRT_FLAG_INLINE_PRIMARY keeps that one).
#define TRT_REF_DONE 0x7FFFFFFEu
#define TRT_PRIM_BURST 8      /* DONE records a lane may consume per S pass (the other lanes of its warp wait meanwhile) */
struct PrimArgs {
    DevScene scene; DevCamera cam; DevParams p;
    float4* a; float4* b;             // [spp_batch][H*W]
    uint32_t tiles_x, tiles_y;
    unsigned long long total;         // tiles * 32 * spp_batch (padding lanes of ragged tiles included)
};
template <int KIND, bool SMEM>
__global__ void __launch_bounds__(256) k_primary(const __grid_constant__ PrimArgs A, const __grid_constant__ FlatList L) {
    const DevScene S = stage_scene<SMEM, KIND>(A.scene);
    const DevParams& P = A.p;
    const unsigned long long per_sample = (unsigned long long)A.tiles_x * A.tiles_y * 32ull, npix = (unsigned long long)P.width * P.height;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.total; i += stride) {
        const uint32_t s = (uint32_t)(i / per_sample);
        const uint32_t r = (uint32_t)(i - (unsigned long long)s * per_sample), tile = r >> 5, ln = r & 31u;
        const uint32_t ty = tile / A.tiles_x, tx = tile - ty * A.tiles_x, x = tx * 8 + (ln & 7u), y = ty * 4 + (ln >> 3);
        if (x >= P.width || y >= P.height) continue;
        const uint32_t pixel = y * P.width + x, sample = P.sample_offset + s;
        V3 o, d;
        primary_ray(A.cam, P, x, y, pixel, sample, &o, &d);
        Hit h;
        if (KIND == 2) h = closest_hit_flat<true>(L.cbox, S.flat_bound, S.flat, S.n_flat, S.flat_qmask[0], S.flat_qmask[1], o, d, P.t_min);
        else h = closest_hit_any<KIND, SMEM, false>(S, o, d, P.t_min, nullptr);
        float4 ra = make_float4(o.x, o.y, o.z, h.t), rb = make_float4(d.x, d.y, d.z, __uint_as_float(h.ref));
        bool over = false; V3 c = mk(0.f, 0.f, 0.f);
        if (h.ref == TRT_REF_INVALID) { over = true; c = mk(P.bg[0], P.bg[1], P.bg[2]); }                  // cpu.rs:58-61
        else {
            Surface sf; hit_material<SMEM>(S, h, &sf);
            if (sf.kind == TRT_MAT_LIGHT) { over = true; c = sf.rgb; }                                      // cpu.rs:49-50, light.rs:17-19
            else if (P.max_bounces <= 1) over = true;                                                       // the scatter would never be traced (cpu.rs:47,54)
        }
        if (over) {
            const V3 col = mk(0.f, 0.f, 0.f) + mk(1.f, 1.f, 1.f) * c;                                       // colour + attenuation * c, as the path loop computes it
            ra = make_float4(col.x, col.y, col.z, h.t); rb.w = __uint_as_float(TRT_REF_DONE);
        }
        const unsigned long long ri = (unsigned long long)s * npix + pixel;
        __stcs(A.a + 2 * ri, ra); __stcs(A.a + 2 * ri + 1, rb);       // one 32-byte record per path: the path loop reads it with ONE 256-bit load
    }
}

// Work fetch for lanes whose item is finished (warp-aggregated atomic). The f64 sums of a pixel owned by one lane are CONTINUED
// from what the accumulation buffer already holds (acc = *dst at fetch, *dst = acc at the end), so that successive launches —
// sample batches of one frame, progressive instalments — sum a pixel's samples in one sequential order, bit for bit as a single
// launch would; a pixel shared by several lanes (tail chunks) adds its partial sums with f64 atomics.
struct ItemState { uint32_t pixel, px, py, s_cur, s_end; double acc0, acc1, acc2; bool has_item, exclusive, done; };
__device__ __forceinline__ void flush_item(const MegaArgs& A, ItemState& I) {
    double* dst = A.sum64 + 3 * (size_t)I.pixel;
    if (I.exclusive) { dst[0] = I.acc0; dst[1] = I.acc1; dst[2] = I.acc2; }
    else { atomicAdd(dst, I.acc0); atomicAdd(dst + 1, I.acc1); atomicAdd(dst + 2, I.acc2); }
    I.has_item = false;
}
__device__ __forceinline__ void fetch_items(const MegaArgs& A, ItemState& I, bool need_item, uint32_t lane, uint32_t* n_paths) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, need_item);
    if (!m) return;
    if (need_item && I.has_item) flush_item(A, I);
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd(A.next_item, (unsigned)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (need_item) {
        const uint32_t item = base + __popc(m & ((1u << lane) - 1u));
        if (item < A.n_items) {
            uint32_t s0, s1;
            if (decode_item(A, item, &I.px, &I.py, &s0, &s1, &I.exclusive)) {
                I.pixel = I.py * A.p.width + I.px; I.s_cur = s0; I.s_end = s1; I.has_item = true;
                if (I.exclusive) { const double* src = A.sum64 + 3 * (size_t)I.pixel; I.acc0 = src[0]; I.acc1 = src[1]; I.acc2 = src[2]; }
                else I.acc0 = I.acc1 = I.acc2 = 0.0;
                if (A.p.max_bounces == 0) { *n_paths += s1 - s0; I.s_cur = s1; }      // no ray is ever cast (cpu.rs:47)
            }   // else: padding item outside the image, fetch again next round
        } else I.done = true;
    }
}
// PRIM: a lane that starts a path takes its first segment from the primary pass. DONE records are added up on the spot (at
// most TRT_PRIM_BURST per pass); the first live record gives (o, d, hit) and the lane goes on to shade it. Returns true if so.
__device__ __forceinline__ bool take_primary(const MegaArgs& A, ItemState& I, V3* o, V3* d, Hit* h, uint32_t* n_rays, uint32_t* n_paths) {
    const size_t npix = (size_t)A.p.width * A.p.height;
#pragma unroll 1
    for (int it = 0; it < TRT_PRIM_BURST && I.s_cur < I.s_end; ++it) {
        const size_t ri = (size_t)I.s_cur * npix + I.pixel;
        // written by k_primary, read-only here. ONE 32-byte record = one sector: with (a, b) in two arrays the final-scene frame read
        // 67 GB from DRAM (profiles/r2j_tree_c2_500spp); interleaved: 131.1 -> 124.4 ms. Prefetching the pixel's next record into L2: +-0.
        const F4x2 rec = ldg256(A.prim_a + 2 * ri);
        const float4 ra = rec.a, rb = rec.b;
        ++*n_rays;
        const uint32_t ref = __float_as_uint(rb.w);
        if (ref != TRT_REF_DONE) { *o = mk(ra.x, ra.y, ra.z); *d = mk(rb.x, rb.y, rb.z); h->t = ra.w; h->ref = ref; return true; }
        I.acc0 += (double)ra.x; I.acc1 += (double)ra.y; I.acc2 += (double)ra.z;          // imager.rs:50, in sample order
        ++I.s_cur; ++*n_paths;
    }
    return false;
}
__device__ __forceinline__ void mega_stats(const MegaArgs& A, uint32_t lane, uint32_t n_rays, uint32_t n_paths, const EventCounters& ec, bool count) {
    unsigned long long r = n_rays, pth = n_paths;
    for (int off = 16; off; off >>= 1) { r += __shfl_xor_sync(0xFFFFFFFFu, r, off); pth += __shfl_xor_sync(0xFFFFFFFFu, pth, off); }
    if (lane == 0) { atomicAdd(A.counters, r); atomicAdd(A.counters + 1, pth); }
    if (count) {
        unsigned long long a = ec.node_visits, b = ec.sphere_tests, c = ec.quad_tests;
        for (int off = 16; off; off >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, off); b += __shfl_xor_sync(0xFFFFFFFFu, b, off); c += __shfl_xor_sync(0xFFFFFFFFu, c, off); }
        if (lane == 0) { atomicAdd(A.counters + 2, a); atomicAdd(A.counters + 3, b); atomicAdd(A.counters + 4, c); }
    }
}

// Loop body = T (one ray segment for every live lane) then S (shade-or-regenerate, fused so that the scattering lanes
// and the lanes starting a new path share ONE Philox call, ONE normalisation and the work-fetch point; every lane that
// still has work leaves S with a live ray, so T always runs with full lanes until the work counter runs dry).
// KIND: 0 = TRT_BVH_REFERENCE, 1 = TRT_BVH_SAH, 2 = TRT_BVH_FLAT (U = the phase-A boxes in the constant bank),
// 3 / 4 = TRT_BVH_SAH / TRT_BVH_REFERENCE walked on their 4-wide nodes. PRIM: paths start from the records of k_primary.
// GEOM_EARLY (flat list only): HitRecord::new for every lane that hit, beside the material lookup (round-1 form), instead of later and
// only for the lanes that scatter. Same operations per lane either way; which placement is faster depends on the scene — with
// spheres in the list (their normal costs a square root and three divisions) early wins (glass Cornell 4K: 252.3 -> 248.4 ms), with
// quads only (precomputed normals) late wins (Cornell 1080p: 52.77 vs 52.93 ms): the launcher picks (profiles/r2l_probe_flat_hit_geometry_placement.log).
template <int KIND, bool SMEM, bool COUNT, bool PRIM, bool GEOM_EARLY = false>
__device__ __forceinline__ void mega_body(const MegaArgs& A, const float4* U) {
    const DevScene S = stage_scene<SMEM, KIND>(A.scene);
    const DevParams& P = A.p;
    const uint32_t lane = threadIdx.x & 31u;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1), atten = mk(1, 1, 1), color = mk(0, 0, 0);
    uint32_t remain = 0, k = 0;
    ItemState I; I.pixel = I.px = I.py = I.s_cur = I.s_end = 0; I.acc0 = I.acc1 = I.acc2 = 0.0; I.has_item = false; I.exclusive = true; I.done = false;
    bool alive = false;
    uint32_t n_rays = 0, n_paths = 0;
    EventCounters ec; ec.node_visits = ec.sphere_tests = ec.quad_tests = 0;

    for (;;) {
        // ---- T: cpu.rs:48
        Hit h; h.t = TRT_INF; h.ref = TRT_REF_INVALID;
        if (alive) {
            if (KIND == 2) h = closest_hit_flat<true>(U, S.flat_bound, S.flat, S.n_flat, S.flat_qmask[0], S.flat_qmask[1], o, d, P.t_min);
            else if (KIND == 3 || KIND == 4) h = closest_hit4<KIND == 3, SMEM, COUNT>(S, o, d, P.t_min, &ec);
            else h = closest_hit<KIND == 1, SMEM, COUNT>(S, o, d, P.t_min, &ec);
            ++n_rays;
        }
        // ---- S1: cpu.rs:49-61 up to the scatter decision: the material alone says "emit and stop" or "scatter"
        Surface sf; sf.kind = TRT_MAT_LIGHT; sf.front = false; sf.param = 0.f;
        sf.point = sf.normal = sf.rgb = mk(0, 0, 0);
        bool scatter = false;
        if (alive) {
            if (h.ref == TRT_REF_INVALID) color = color + atten * mk(P.bg[0], P.bg[1], P.bg[2]);
            else {
                if (KIND == 2 && !PRIM && GEOM_EARLY) hit_geometry<SMEM>(S, h, o, d, &sf);
                hit_material<SMEM>(S, h, &sf);
                V3 emission = (sf.kind == TRT_MAT_LIGHT) ? sf.rgb : mk(0.0f, 0.0f, 0.0f);
                color = color + atten * emission;
                // a scattered ray is only traced while bounces remain (cpu.rs:47,54): the last scatter is never used
                scatter = sf.kind != TRT_MAT_LIGHT && remain > 1;
            }
            if (!scatter) {       // imager.rs:50 (f64 here)
                I.acc0 += (double)color.x; I.acc1 += (double)color.y; I.acc2 += (double)color.z;
                alive = false; ++I.s_cur; ++n_paths;
            }
        }
        // ---- S2: work fetch for lanes whose chunk is finished
        fetch_items(A, I, !alive && !I.done && I.s_cur == I.s_end, lane, &n_paths);
        // ---- S3: the next ray of every lane that scatters (material/*.rs) or starts a new path (pointgen.rs:37-52)
        const bool regen = !alive && !I.done && I.s_cur < I.s_end;
        if (PRIM) {
            bool fresh = false;
            if (regen) fresh = take_primary(A, I, &o, &d, &h, &n_rays, &n_paths);
            if (fresh) { hit_material<SMEM>(S, h, &sf); atten = mk(1, 1, 1); color = mk(0, 0, 0); remain = P.max_bounces; k = 0; alive = true; }
            if (scatter || fresh) {
                hit_geometry<SMEM>(S, h, o, d, &sf);
                ++k;
                U4 r = philox4x32_10(I.pixel, P.sample_offset + I.s_cur, k, 0u, P.seed_lo, P.seed_hi);
                V3 dir = scatter_direction(sf, d, r);
                atten = atten * sf.rgb; o = sf.point; --remain;                               // cpu.rs:52-54
                d = normalized(dir);                                                         // Ray::new, ray.rs:12-14
            }
        } else {
            if (scatter && !(KIND == 2 && GEOM_EARLY)) hit_geometry<SMEM>(S, h, o, d, &sf);
            if (scatter || regen) {
                uint32_t sample = P.sample_offset + I.s_cur;
                if (scatter) ++k; else k = 0;
                U4 r = philox4x32_10(I.pixel, sample, k, 0u, P.seed_lo, P.seed_hi);
                V3 dir;
                if (scatter) {
                    dir = scatter_direction(sf, d, r);
                    atten = atten * sf.rgb; o = sf.point; --remain;                           // cpu.rs:52-54
                } else {
                    camera_ray(A.cam, P, I.px, I.py, I.pixel, sample, r, &o, &dir);
                    atten = mk(1, 1, 1); color = mk(0, 0, 0); remain = P.max_bounces; alive = true;
                }
                d = normalized(dir);                                                         // Ray::new, ray.rs:12-14
            }
        }
        if (__all_sync(0xFFFFFFFFu, I.done)) break;
    }
    mega_stats(A, lane, n_rays, n_paths, ec, COUNT);
}

// The tree kinds use this form of the same loop: the traversal is advanced ONE step (trav_step) per iteration, and the
// S block runs as soon as `A.regen_min` lanes have finished their ray (or nobody is traversing any more). With
// regen_min = 32 this is the T-then-S schedule above; smaller values stop a warp from idling behind its slowest ray,
// which pays when a ray costs many steps (1 M spheres: ~50 node steps per ray and 7.7 of 32 lanes active under the
// T-then-S schedule, profiles/r1g) and costs when S (≈ 800 instructions) would run for a handful of lanes on a
// 7-step Cornell ray. The launcher picks regen_min from the tree size; trt_params.flags bits 8..15 override it.
// STEP (wide nodes in global memory): 1 = trav4_step with 128-bit node loads (round 1), 2 = trav4_step with 256-bit node loads,
// 3 = trav4_step_pipe (one dependent fetch per step, 256-bit node loads). Elsewhere STEP is 1 and means trav4_step / trav_step.
template <int KIND, bool SMEM, bool COUNT, bool PRIM, int STEP>
__device__ __forceinline__ void mega_body_async(const MegaArgs& A) {
    const DevScene S = stage_scene<SMEM, KIND>(A.scene);
    const DevParams& P = A.p;
    const uint32_t lane = threadIdx.x & 31u;
    Trav T; T.cur = TRT_REF_INVALID; T.sp = 0; T.best.t = TRT_INF; T.best.ref = TRT_REF_INVALID;
    T.o = mk(0, 0, 0); T.d = mk(0, 0, 1); T.inv = mk(0, 0, 1); T.t_min = P.t_min; T.tame = false; T.e8 = 0.f; T.replay = false; T.best_anom = false; T.cur_s = 0.f;
    TravStack K;
    V3 atten = mk(1, 1, 1), color = mk(0, 0, 0);
    uint32_t remain = 0, k = 0;
    ItemState I; I.pixel = I.px = I.py = I.s_cur = I.s_end = 0; I.acc0 = I.acc1 = I.acc2 = 0.0; I.has_item = false; I.exclusive = true; I.done = false;
    bool alive = false;      // alive: the lane carries a path; T.cur != INVALID: its ray is still being traced
    uint32_t n_rays = 0, n_paths = 0;
    EventCounters ec; ec.node_visits = ec.sphere_tests = ec.quad_tests = 0;
    const int regen_min = (int)A.regen_min;

    for (;;) {
        // ---- T: one step of cpu.rs:48 for every lane whose ray is in flight
        if (T.cur != TRT_REF_INVALID) {
            if (KIND == 5) {            // compressed 8-wide nodes (ordered; global memory)
                trav8_any_step<COUNT>(T, S, P.t_min, K, &ec);
                if (T.cur == TRT_REF_INVALID && needs_replay(T, S)) begin_replay(T, S);                    // rare: DESIGN.md §4.3
            }
            else if (KIND == 3 || KIND == 4) {
                if (STEP == 3 && !SMEM) trav4_any_step_pipe<KIND == 3, COUNT>(T, S, P.t_min, K, &ec);
                else trav4_any_step<KIND == 3, SMEM, COUNT, STEP == 2>(T, S, P.t_min, K, &ec);
                if (KIND == 3 && T.cur == TRT_REF_INVALID && needs_replay(T, S)) begin_replay(T, S);     // rare: DESIGN.md §4.3
            }
            else trav_step<KIND == 1, SMEM, COUNT>(T, S, K, &ec);
        }
        const bool tracing = T.cur != TRT_REF_INVALID;
        const unsigned tm = __ballot_sync(0xFFFFFFFFu, tracing);
        const unsigned wm = __ballot_sync(0xFFFFFFFFu, !tracing && !I.done);
        if (tm != 0u && __popc(wm) < regen_min) continue;
        // ---- S1: cpu.rs:49-61 up to the scatter decision, for the lanes whose ray just finished
        const bool in_s = !tracing && !I.done;
        Surface sf; sf.kind = TRT_MAT_LIGHT; sf.front = false; sf.param = 0.f;
        sf.point = sf.normal = sf.rgb = mk(0, 0, 0);
        bool scatter = false;
        Hit h = T.best; V3 ro = T.o, rd = T.d;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_s && alive) {
            if (KIND == 1 && needs_replay(T, S)) h = ordered_replay(S, T.o, T.d, P.t_min);   // rare: DESIGN.md §4.3 (binary walk)
            if (h.ref == TRT_REF_INVALID) color = color + atten * mk(P.bg[0], P.bg[1], P.bg[2]);
            else {
                if (!SMEM && !PRIM) g0 = hit_prim_f4<SMEM>(S, h);      // global memory: fetched beside the meta record, not after the material
                hit_material<SMEM>(S, h, &sf);
                V3 emission = (sf.kind == TRT_MAT_LIGHT) ? sf.rgb : mk(0.0f, 0.0f, 0.0f);
                color = color + atten * emission;
                scatter = sf.kind != TRT_MAT_LIGHT && remain > 1;
            }
            if (!scatter) {       // imager.rs:50 (f64 here)
                I.acc0 += (double)color.x; I.acc1 += (double)color.y; I.acc2 += (double)color.z;
                alive = false; ++I.s_cur; ++n_paths;
            }
        }
        // ---- S2: work fetch for lanes whose chunk is finished
        fetch_items(A, I, in_s && !alive && I.s_cur == I.s_end, lane, &n_paths);
        // ---- S3: next ray of the lanes that scatter or start a new path; its traversal state is set up here
        const bool regen = in_s && !alive && !I.done && I.s_cur < I.s_end;
        if (PRIM) {
            bool fresh = false;
            if (regen) fresh = take_primary(A, I, &ro, &rd, &h, &n_rays, &n_paths);
            if (fresh) { hit_material<SMEM>(S, h, &sf); atten = mk(1, 1, 1); color = mk(0, 0, 0); remain = P.max_bounces; k = 0; alive = true; }
            if (scatter || fresh) {
                hit_geometry<SMEM>(S, h, ro, rd, &sf);
                ++k;
                U4 r = philox4x32_10(I.pixel, P.sample_offset + I.s_cur, k, 0u, P.seed_lo, P.seed_hi);
                V3 dir = scatter_direction(sf, rd, r);
                atten = atten * sf.rgb; --remain;
                trav_begin<COUNT && (KIND < 3)>(T, S, sf.point, normalized(dir), P.t_min, &ec);
                if (KIND == 5) trav8_setup(T, S);
                ++n_rays;
            }
        } else {
            if (scatter) { if (!SMEM) hit_geometry_f4(S, h, ro, rd, g0, &sf); else hit_geometry<SMEM>(S, h, ro, rd, &sf); }
            if (scatter || regen) {
                uint32_t sample = P.sample_offset + I.s_cur;
                if (scatter) ++k; else k = 0;
                U4 r = philox4x32_10(I.pixel, sample, k, 0u, P.seed_lo, P.seed_hi);
                V3 o, dir;
                if (scatter) {
                    dir = scatter_direction(sf, rd, r);
                    atten = atten * sf.rgb; o = sf.point; --remain;
                } else {
                    camera_ray(A.cam, P, I.px, I.py, I.pixel, sample, r, &o, &dir);
                    atten = mk(1, 1, 1); color = mk(0, 0, 0); remain = P.max_bounces; alive = true;
                }
                trav_begin<COUNT && (KIND < 3)>(T, S, o, normalized(dir), P.t_min, &ec);      // Ray::new normalises, ray.rs:12-14 (root-box event only for the binary counts)
                if (KIND == 5) trav8_setup(T, S);
                ++n_rays;
                // an empty scene (or, when counting, a ray that misses the root box) is finished at once: T.cur stays INVALID
                // and the lane re-enters S on the next pass with T.best = miss.
            }
        }
        if (__all_sync(0xFFFFFFFFu, I.done)) break;
    }
    mega_stats(A, lane, n_rays, n_paths, ec, COUNT);
}

// MODE: 0 = T-then-S (small trees), 1 = step-wise (big trees). Separate kernels, so that each loop gets its own register
// allocation (one kernel holding both loops spilled 244 bytes in the step-wise loop, its own kernel 94).
// Measured and rejected (round 2): a step-wise loop with TWO paths per lane whose state lives in shared memory, so that a lane
// traces its second path while the hit of the first waits for the S block: correct (bit-identical sums, in sample order), but
// 10^6 spheres 1 666 vs 2 128 Mrays/s and the final scene 4 209 vs 7 710 — the 35 KB of shared memory per CTA it needs cost
// L1 capacity (10^6 spheres) or a whole CTA of occupancy (final scene), and the ray-swap block runs at low lane counts.
template <int KIND, bool SMEM, bool COUNT, int MODE, bool PRIM>
__global__ void __launch_bounds__(MEGA_THREADS, TRT_MEGA_MIN_BLOCKS) k_render_mega(const __grid_constant__ MegaArgs A) {
    if constexpr (MODE == 0) mega_body<KIND, SMEM, COUNT, PRIM>(A, nullptr);
    else mega_body_async<KIND, SMEM, COUNT, PRIM, MODE>(A);
}
template <bool PRIM, bool GEOM_EARLY = false>
__global__ void __launch_bounds__(TRT_FLAT_THREADS, TRT_FLAT_MIN_BLOCKS) k_render_mega_flat(const __grid_constant__ MegaArgs A, const __grid_constant__ FlatList L) {
    mega_body<2, true, false, PRIM, GEOM_EARLY>(A, L.cbox);
}

// ------------------------------------------------------------------ AOV / sample batch / trace
struct AovArgs {
    DevScene scene; DevCamera cam; DevParams p;
    uint32_t* prim_id; float* t; float* radiance; uint32_t* n_rays;
    unsigned long long total;   // W*H*spp
};
__device__ __forceinline__ uint32_t prim_id_of(const DevScene& S, uint32_t ref) {
    if (ref == TRT_REF_INVALID) return 0xFFFFFFFFu;
    uint32_t idx = ref & TRT_REF_INDEX_MASK;
    return (ref & TRT_REF_QUAD) ? S.quad_meta[idx].z : S.sphere_meta[idx].z;
}
template <int KIND>
__global__ void __launch_bounds__(128) k_aov(const AovArgs A) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.total) return;
    const DevScene& S = A.scene; const DevParams& P = A.p;
    uint32_t pixel = (uint32_t)(i / P.spp), si = (uint32_t)(i - (unsigned long long)pixel * P.spp);
    uint32_t y = pixel / P.width, x = pixel - y * P.width, sample = P.sample_offset + si;
    V3 o, d;
    primary_ray(A.cam, P, x, y, pixel, sample, &o, &d);
    DevParams Q = P;
    if (!A.radiance && !A.n_rays) Q.max_bounces = P.max_bounces ? 1u : 0u;
    Hit first; first.t = TRT_INF; first.ref = TRT_REF_INVALID;
    uint32_t nr = 0; EventCounters ec;
    V3 col = trace_path<KIND, false, false>(S, Q, o, d, pixel, sample, &nr, &first, &ec);
    if (A.prim_id) A.prim_id[i] = prim_id_of(S, first.ref);
    if (A.t) A.t[i] = first.t;
    if (A.radiance) { A.radiance[3 * i] = col.x; A.radiance[3 * i + 1] = col.y; A.radiance[3 * i + 2] = col.z; }
    if (A.n_rays) A.n_rays[i] = nr;
}

struct BatchArgs { DevScene scene; DevParams p; const trt_sample_point_pod* in; trt_sampled_color_pod* out; unsigned long long n; };
template <int KIND>
__global__ void __launch_bounds__(128) k_sample_batch(const BatchArgs A) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    trt_sample_point_pod sp = A.in[i];
    uint32_t pixel = sp.y * A.p.width + sp.x;
    EventCounters ec;
    V3 col = trace_path<KIND, false, false>(A.scene, A.p, mk(sp.o[0], sp.o[1], sp.o[2]), mk(sp.d[0], sp.d[1], sp.d[2]), pixel,
                                            A.p.sample_offset + (uint32_t)i, nullptr, nullptr, &ec);
    trt_sampled_color_pod c; c.x = sp.x; c.y = sp.y; c.c[0] = col.x; c.c[1] = col.y; c.c[2] = col.z;
    A.out[i] = c;
}

struct TraceArgs { DevScene scene; const float* o; const float* d; float t_min; uint32_t* prim_id; float* t; unsigned long long n; };
template <int KIND>
__global__ void __launch_bounds__(128) k_trace(const TraceArgs A) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    EventCounters ec;
    Hit h = closest_hit_any<KIND, false, false>(A.scene, mk(A.o[3 * i], A.o[3 * i + 1], A.o[3 * i + 2]),
                                                mk(A.d[3 * i], A.d[3 * i + 1], A.d[3 * i + 2]), A.t_min, &ec);
    A.prim_id[i] = prim_id_of(A.scene, h.ref); A.t[i] = h.t;
}

// ------------------------------------------------------------------ epilogue
__global__ void __launch_bounds__(256) k_finalize(const double* __restrict__ sum64, size_t n, double inv_spp,
                                                  float* __restrict__ linear, uint8_t* __restrict__ rgb8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lin = (float)(sum64[i] * inv_spp);
    if (linear) linear[i] = lin;
    if (rgb8) rgb8[i] = tonemap_u8(lin);
}

__global__ void k_debug_math(size_t n, const float* u, float* s, float* c, float* cb) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ss, cc; det_sincos2pi(u[i], &ss, &cc);
    s[i] = ss; c[i] = cc; cb[i] = det_cbrt(u[i]);
}
__global__ void k_debug_philox(size_t n, const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    U4 r = philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

// ------------------------------------------------------------------ launchers
namespace trtk {

uint32_t walk_kind(const SceneView& v, bool binary_only) {
    if (binary_only || !v.nodes4 || !v.n_nodes4) return v.kind;
    return v.kind == 1 ? 3u : (v.kind == 0 ? 4u : v.kind);
}
// kind: as walked (3 = the 4-wide nodes of a TRT_BVH_SAH scene)
size_t scene_smem_bytes(const SceneView& v, uint32_t kind) {
    size_t aux = aux_f4(v.n_spheres, v.n_quads, v.n_materials);
    if (aux > TRT_AUX_SMEM_F4) aux = 0;
    size_t nodes = (kind == 3 || kind == 4) ? (size_t)TRT_WIDE_SMEM_F4 * v.n_nodes4 : 4ull * v.n_nodes;
    return 16ull * (nodes + v.n_spheres + 4ull * v.n_quads + (size_t)TRT_FLAT_F4 * v.n_flat + aux);
}

template <int KIND, bool SMEM, bool COUNT, int MODE, bool PRIM>
static cudaError_t mega_fn4(const void** fn, int* blocks_per_sm, size_t smem) {
    auto kern = k_render_mega<KIND, SMEM, COUNT, MODE, PRIM>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // big scenes stage nothing: give the whole configurable L1 / shared-memory array to L1 (their node and stack traffic lives there)
    if (!SMEM) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxL1); if (e != cudaSuccess) return e; }
    *fn = (const void*)kern;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, MEGA_THREADS, smem);
}
// mode: 0 T-then-S, 1 step-wise, 2 / 3 step-wise with the other step functions (wide nodes in global memory, no event counting: else 1);
// prim: paths start from k_primary's records (never together with event counting)
template <int KIND, bool SMEM, bool COUNT>
static cudaError_t mega_fn(int mode, bool prim, const void** fn, int* blocks_per_sm, size_t smem) {
    if constexpr ((KIND == 3 || KIND == 4) && !SMEM && !COUNT) {
        if (mode == 2) return prim ? mega_fn4<KIND, false, false, 2, true>(fn, blocks_per_sm, smem) : mega_fn4<KIND, false, false, 2, false>(fn, blocks_per_sm, smem);
        if (mode == 3) return prim ? mega_fn4<KIND, false, false, 3, true>(fn, blocks_per_sm, smem) : mega_fn4<KIND, false, false, 3, false>(fn, blocks_per_sm, smem);
    }
    if (mode > 1) mode = 1;
    if (prim && !COUNT) return mode == 1 ? mega_fn4<KIND, SMEM, false, 1, true>(fn, blocks_per_sm, smem) : mega_fn4<KIND, SMEM, false, 0, true>(fn, blocks_per_sm, smem);
    return mode == 1 ? mega_fn4<KIND, SMEM, COUNT, 1, false>(fn, blocks_per_sm, smem) : mega_fn4<KIND, SMEM, COUNT, 0, false>(fn, blocks_per_sm, smem);
}

template <int KIND, bool SMEM>
static cudaError_t primary_launch(const PrimArgs& A, const FlatList& L, size_t smem, int sm_count, cudaStream_t st) {
    auto kern = k_primary<KIND, SMEM>;
    cudaError_t e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    int bps = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, smem)) != cudaSuccess) return e;
    if (bps < 1) bps = 1;
    unsigned long long want = (A.total + 255) / 256, cap = (unsigned long long)sm_count * bps * 8;     // a few waves of persistent CTAs
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    kern<<<grid ? grid : 1, 256, smem, st>>>(A, L);
    return cudaGetLastError();
}

size_t primary_record_bytes(const trt_params& p) { return (size_t)p.width * p.height * 32ull; }    // per sample of the batch
// Measured on B200 (64-spp probes, profiles/r2f_*): final scene (486 wide nodes in shared memory, T-then-S) 7 917 -> 8 476 Mrays/s
// with the pass; Cornell flat list 22 299 -> 19 254, glass Cornell 13 007 -> 12 432, 10^6 spheres (L2, step-wise) 2 119 -> 2 089:
// there a primary ray is just one more ray of the warp's pool and the records' HBM round trip costs more than it saves.
bool primary_pass_pays(const SceneView& v, const trt_params& p) {
    if (v.kind == 2) return false;
    const uint32_t kind = walk_kind(v, (p.flags & TRT_FLAG_BINARY_NODES) != 0);
    if (kind != 3 && kind != 4) return false;
    if (scene_smem_bytes(v, kind) > 64 * 1024) return false;
    const uint32_t regen = (p.flags >> 8) & 0xFFu;
    return regen == 0 ? v.n_nodes <= 4096 : regen >= 32u;             // the T-then-S loop
}

cudaError_t launch_render_mega(const SceneView& v, const trt_camera& cam, const trt_params& p, double* d_sum64,
                               unsigned long long* d_counters, unsigned int* d_next_item, int sm_count, cudaStream_t st,
                               void* d_primary) {
    MegaArgs A;
    A.scene = dev_scene(v); A.cam = dev_camera(cam); A.p = dev_params(p);
    A.sum64 = d_sum64; A.counters = d_counters; A.next_item = d_next_item;
    // d_primary: room for primary_record_bytes(p) * p.spp -> the coherent first-bounce pass runs first and the path loop starts
    // from its records (never when events are counted or no ray is cast at all)
    const bool prim = d_primary != nullptr && p.spp > 0 && p.max_bounces > 0 && !(p.flags & (1u | TRT_FLAG_COUNT_WALKED | TRT_FLAG_INLINE_PRIMARY));
    const size_t npix = (size_t)p.width * p.height;
    A.prim_a = prim ? (const float4*)d_primary : nullptr;
    A.prim_b = prim ? (const float4*)d_primary + npix * p.spp : nullptr;
    // shade/regenerate threshold: trees of a few dozen nodes cost ~7 steps per ray and S is ~800 instructions, so S must run
    // for the whole warp; big trees cost ~50 steps per ray and waiting for the slowest of 32 rays wastes 3/4 of the lanes.
    uint32_t regen = (p.flags >> 8) & 0xFFu;
    // measured (B200, regen sweep): Cornell 17 nodes, glass Cornell 33, final scene 486: 32 best; 1 M spheres: 16 best (+25 %)
    if (regen == 0) regen = v.n_nodes <= 4096 ? 32u : 16u;       // (re-set below for 4-wide nodes, whose rays take half as many steps: 20)
    A.regen_min = regen > 32u ? 32u : regen;
    const bool count_walked = (p.flags & TRT_FLAG_COUNT_WALKED) != 0 && !(p.flags & 1u);
    bool count = (p.flags & 1u) != 0;
    // a TRT_BVH_SAH scene is walked on its 4-wide nodes unless events are counted (TRT_FLAG_COUNT_EVENTS' counts are defined on
    // binary nodes; TRT_FLAG_COUNT_WALKED counts the records of whatever walk the scene would get anyway)
    const uint32_t kind = walk_kind(v, count || (p.flags & TRT_FLAG_BINARY_NODES));
    if (count_walked) count = true;
    if ((kind == 3 || kind == 4) && ((p.flags >> 8) & 0xFFu) == 0 && v.n_nodes > 4096) A.regen_min = 20u;    // measured: 1 M spheres 2 315 (20) vs 2 282 (16) Mrays/s
    size_t smem = scene_smem_bytes(v, kind);
    bool use_smem = smem <= 64 * 1024;
    if (v.kind == 2 && (!use_smem || v.n_flat > TRT_FLAT_MAX || !v.flat_cbox_host)) return cudaErrorInvalidValue;
    if (!use_smem) smem = 0;
    // step-wise loop: which step function (flags bits 16..17: 0 = chosen here, 1 = 128-bit node loads, 2 = 256-bit, 3 = 256-bit and one fetch per step)
    const uint32_t step_req = (p.flags >> 16) & 3u;
    const int mega_mode = A.regen_min >= 32u ? 0 : (step_req ? (int)step_req : TRT_STEP_DEFAULT);
    const void* fn = nullptr; int blocks_per_sm = 0; cudaError_t e;
    // ordered trees in global memory, step-wise loop: the compressed 8-wide nodes when the scene has them (trt_scene_upload builds
    // them only on request, TRT_WIDE8=1: measured slower than the 4-wide walk on B200); TRT_FLAG_NO_WIDE8 keeps the 4-wide walk
    // even then (same results; A/B, tests)
    const bool use8 = kind == 3 && !use_smem && mega_mode != 0 && v.nodes8 && v.n_nodes8 && v.nodes_ref4 && !prim && !(p.flags & (1u | TRT_FLAG_NO_WIDE8));
    if (use8) {
        e = count ? mega_fn4<5, false, true, 2, false>(&fn, &blocks_per_sm, smem) : mega_fn4<5, false, false, 2, false>(&fn, &blocks_per_sm, smem);
    } else if (v.kind == 2) {
        fn = prim ? (const void*)k_render_mega_flat<true> : (v.n_spheres > 0 ? (const void*)k_render_mega_flat<false, true> : (const void*)k_render_mega_flat<false, false>);
        if (smem > 48 * 1024) { e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, TRT_FLAT_THREADS, smem);
    } else if (kind == 3) {
        if (count) e = use_smem ? mega_fn<3, true, true>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<3, false, true>(mega_mode, prim, &fn, &blocks_per_sm, smem);
        else e = use_smem ? mega_fn<3, true, false>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<3, false, false>(mega_mode, prim, &fn, &blocks_per_sm, smem);
    } else if (kind == 4) {
        if (count) e = use_smem ? mega_fn<4, true, true>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<4, false, true>(mega_mode, prim, &fn, &blocks_per_sm, smem);
        else e = use_smem ? mega_fn<4, true, false>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<4, false, false>(mega_mode, prim, &fn, &blocks_per_sm, smem);
    } else if (v.kind == 1) {
        if (use_smem) e = count ? mega_fn<1, true, true>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<1, true, false>(mega_mode, prim, &fn, &blocks_per_sm, smem);
        else e = count ? mega_fn<1, false, true>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<1, false, false>(mega_mode, prim, &fn, &blocks_per_sm, smem);
    } else {
        if (use_smem) e = count ? mega_fn<0, true, true>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<0, true, false>(mega_mode, prim, &fn, &blocks_per_sm, smem);
        else e = count ? mega_fn<0, false, true>(mega_mode, prim, &fn, &blocks_per_sm, smem) : mega_fn<0, false, false>(mega_mode, prim, &fn, &blocks_per_sm, smem);
    }
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    // persistent grid: every SM filled to the kernel's occupancy limit, one wave
    int grid = sm_count * blocks_per_sm;
    const unsigned threads = v.kind == 2 ? TRT_FLAT_THREADS : MEGA_THREADS;
    uint64_t resident = (uint64_t)grid * threads;
    uint64_t pixels = (uint64_t)p.width * p.height;
    // work items: whole pixels (all samples, plain store) while there are plenty of them; the LAST tiles of the list are
    // split into chunks of samples so that no lane idles behind one long final pixel (at 1024 spp a pixel is 1/14 of a
    // lane's whole job). With few pixels everything is split so that the lanes see >= ~8 items each. Never < 4 samples.
    uint32_t chunk = p.spp ? p.spp : 1;
    A.tiles_x = (p.width + 7) / 8; A.tiles_y = (p.height + 3) / 4;
    uint64_t n_tiles = (uint64_t)A.tiles_x * A.tiles_y, tiles_rest = n_tiles;
    if (pixels < 8 * resident) {
        uint64_t want_chunks = (8 * resident + pixels - 1) / (pixels ? pixels : 1);
        uint64_t c = (p.spp + want_chunks - 1) / want_chunks;
        if (c < 4) c = 4;
        if (c < chunk) chunk = (uint32_t)c;
    } else if (p.spp >= 32 && !(p.flags & TRT_FLAG_NO_TAIL_SPLIT)) {
        // tuning overrides (development): TRT_TAIL_PX = pixels per lane handed out in pieces, TRT_TAIL_DIV = pieces per pixel, TRT_TAIL_MIN = samples per piece at least
        double tail_px = 3.0; uint32_t tail_div = 32, tail_min = 4;      // measured (Cornell 1080p, 128 / 256 spp, profiles/r2j_probe_c3_tail_split.log): 3 pixels per lane -0.4 % against 2; 0.5...1.5 slower; 16...128 pieces +-0.2 %
        if (const char* e = std::getenv("TRT_TAIL_PX")) { double x = std::atof(e); if (x >= 0.0 && x <= 64.0) tail_px = x; }
        if (const char* e = std::getenv("TRT_TAIL_DIV")) { long x = std::atol(e); if (x >= 1 && x <= 4096) tail_div = (uint32_t)x; }
        if (const char* e = std::getenv("TRT_TAIL_MIN")) { long x = std::atol(e); if (x >= 1 && x <= 4096) tail_min = (uint32_t)x; }
        tiles_rest = (uint64_t)((tail_px * (double)resident + 31.0) / 32.0);                 // ~3 pixels per lane are handed out in pieces
        if (tiles_rest > n_tiles) tiles_rest = n_tiles;
        chunk = (p.spp + tail_div - 1) / tail_div; if (chunk < tail_min) chunk = tail_min;     // 1/32 of a pixel: the job ends within ~half a chunk per lane (1/8: -0.35 % at 1024 spp)
        if (chunk > p.spp) chunk = p.spp;
    }
    A.chunk = chunk; A.n_chunks = p.spp ? (p.spp + chunk - 1) / chunk : 0;
    A.tiles_whole = (uint32_t)(n_tiles - tiles_rest);
    uint64_t n_items = 32ull * (A.tiles_whole + tiles_rest * A.n_chunks);
    if (p.spp == 0) n_items = 0;
    if (n_items > 0xFFFFFFF0ull - resident) return cudaErrorInvalidValue;
    A.n_items = (uint32_t)n_items;
    FlatList L;
    if (v.kind == 2) memcpy(&L, v.flat_cbox_host, sizeof L); else memset(&L, 0, sizeof L);
    if (prim) {
        PrimArgs PA;
        PA.scene = A.scene; PA.cam = A.cam; PA.p = A.p; PA.a = (float4*)d_primary; PA.b = (float4*)d_primary + npix * p.spp;
        PA.tiles_x = A.tiles_x; PA.tiles_y = A.tiles_y; PA.total = n_tiles * 32ull * p.spp;
        if (v.kind == 2) e = primary_launch<2, true>(PA, L, smem, sm_count, st);
        else if (kind == 3) e = use_smem ? primary_launch<3, true>(PA, L, smem, sm_count, st) : primary_launch<3, false>(PA, L, smem, sm_count, st);
        else if (kind == 4) e = use_smem ? primary_launch<4, true>(PA, L, smem, sm_count, st) : primary_launch<4, false>(PA, L, smem, sm_count, st);
        else if (v.kind == 1) e = use_smem ? primary_launch<1, true>(PA, L, smem, sm_count, st) : primary_launch<1, false>(PA, L, smem, sm_count, st);
        else e = use_smem ? primary_launch<0, true>(PA, L, smem, sm_count, st) : primary_launch<0, false>(PA, L, smem, sm_count, st);
        if (e != cudaSuccess) return e;
    }
    if (v.kind == 2) {
        void* args[] = {(void*)&A, (void*)&L};
        return cudaLaunchKernel(fn, dim3(grid), dim3(threads), args, smem, st);
    }
    void* args[] = {(void*)&A};
    return cudaLaunchKernel(fn, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}

// the one-thread-per-item kernels walk the compressed 8-wide nodes whenever the scene has them (never staged: global memory)
static uint32_t item_walk_kind(const SceneView& v) {
    const uint32_t k = walk_kind(v, false);
    return (k == 3 && v.nodes8 && v.n_nodes8 && v.nodes_ref4) ? 5u : k;
}
#define TRT_KIND_LAUNCH(kern, kind, blocks, threads, st, A) \
    do { if ((kind) == 5) kern<5><<<(blocks), (threads), 0, (st)>>>(A); else if ((kind) == 2) kern<2><<<(blocks), (threads), 0, (st)>>>(A); else if ((kind) == 1) kern<1><<<(blocks), (threads), 0, (st)>>>(A); \
         else if ((kind) == 3) kern<3><<<(blocks), (threads), 0, (st)>>>(A); else if ((kind) == 4) kern<4><<<(blocks), (threads), 0, (st)>>>(A); else kern<0><<<(blocks), (threads), 0, (st)>>>(A); } while (0)

cudaError_t launch_aov(const SceneView& v, const trt_camera& cam, const trt_params& p, uint32_t* d_prim, float* d_t,
                       float* d_rad, uint32_t* d_nrays, cudaStream_t st) {
    AovArgs A; A.scene = dev_scene(v); A.cam = dev_camera(cam); A.p = dev_params(p);
    A.prim_id = d_prim; A.t = d_t; A.radiance = d_rad; A.n_rays = d_nrays;
    A.total = (unsigned long long)p.width * p.height * p.spp;
    if (A.total == 0) return cudaSuccess;
    unsigned long long blocks = (A.total + 127) / 128;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    TRT_KIND_LAUNCH(k_aov, item_walk_kind(v), (unsigned)blocks, 128, st, A);
    return cudaGetLastError();
}

cudaError_t launch_sample_batch(const SceneView& v, const trt_params& p, const trt_sample_point_pod* d_in, size_t n,
                                trt_sampled_color_pod* d_out, cudaStream_t st) {
    BatchArgs A; A.scene = dev_scene(v); A.p = dev_params(p); A.in = d_in; A.out = d_out; A.n = n;
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n + 127) / 128;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    TRT_KIND_LAUNCH(k_sample_batch, item_walk_kind(v), (unsigned)blocks, 128, st, A);
    return cudaGetLastError();
}

cudaError_t launch_trace(const SceneView& v, const float* d_o, const float* d_d, size_t n, float t_min, uint32_t* d_prim,
                         float* d_t, cudaStream_t st) {
    TraceArgs A; A.scene = dev_scene(v); A.o = d_o; A.d = d_d; A.t_min = t_min; A.prim_id = d_prim; A.t = d_t; A.n = n;
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n + 127) / 128;
    if (blocks > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    TRT_KIND_LAUNCH(k_trace, item_walk_kind(v), (unsigned)blocks, 128, st, A);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const double* d_sum64, size_t n_values, uint32_t spp_total, float* d_linear, uint8_t* d_rgb8, cudaStream_t st) {
    if (n_values == 0) return cudaSuccess;
    double inv = 1.0 / (double)spp_total;
    size_t blocks = (n_values + 255) / 256;
    k_finalize<<<(unsigned)blocks, 256, 0, st>>>(d_sum64, n_values, inv, d_linear, d_rgb8);
    return cudaGetLastError();
}

cudaError_t launch_debug_math(size_t n, const float* u, float* s, float* c, float* cb, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_debug_math<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, u, s, c, cb);
    return cudaGetLastError();
}
cudaError_t launch_debug_philox(size_t n, const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    k_debug_philox<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ctr, k0, k1, out);
    return cudaGetLastError();
}

} // namespace trtk
