// trt_wavefront.cu — the wavefront variant (TRT_VARIANT_WAVEFRONT) of the path-tracing hot path for sm_100a.
//
// A fixed pool of path slots lives in HBM as struct-of-float4 arrays (sized to stay L2-resident); two index queues
// hold the slots that carry a ray. One bounce of the whole pool is two kernels:
//
//   k_wf_shade   one thread per queued slot: finishes the previous segment (cpu.rs:49-61: emission, miss -> background,
//                scatter decision), adds finished paths to the f64 pixel accumulators, refills a finished slot with the
//                next camera path while any remain (pointgen.rs:37-52), scatters the others (material/*.rs), and
//                appends every slot that still carries a ray to the next queue by warp ballot + popc + one atomic per warp.
//   k_wf_extend  persistent warps over the queue: closest_hit in its resumable form (trav_begin / trav_step); a lane
//                whose ray is finished writes (t, ref) back and immediately takes the next queued ray, so a warp never
//                waits for its slowest ray — the traversal-divergence cure the megakernel cannot apply, because there a
//                finished lane needs ~500 instructions of shading before it has another ray.
//
// Path p of the job maps to (8x4 pixel tile, sample, lane) so that 32 consecutive paths are the 32 pixels of one tile at
// the same sample index; the Philox streams are keyed by (pixel, global sample, bounce) exactly as in the megakernel,
// so both variants trace the same paths and differ only in the order of the f64 accumulation.
#include "trt_device.cuh"
#include "trt_kernels.h"
#include "trt_host_glue.cuh"

using namespace trtd;

#define WF_THREADS 256
#define WF_REFILL_MIN 1      // refill idle lanes as soon as this many are idle

struct WfCtl {                         // device-resident control block
    unsigned int q_count[2];           // entries in queue 0 / 1
    unsigned int q_head;               // extend: next queue entry to hand out
    unsigned int pad;
    unsigned long long next_path;      // shade: next path index of the job
    unsigned long long rays, paths;    // statistics
};

struct WfArgs {
    DevScene scene; DevCamera cam; DevParams p;
    float4* a0;        // (o.xyz, d.x)
    float4* a1;        // (d.y, d.z, hit t, hit ref bits)
    float4* a2;        // (attenuation rgb, pixel bits)
    float4* a3;        // (colour rgb, sample bits)
    uint32_t* rk;      // remain | k << 16 ; 0xFFFFFFFF = slot carries no path
    uint32_t* queue[2];
    WfCtl* ctl;
    double* sum64;
    unsigned long long total_paths;    // tiles * 32 * spp (padding paths of ragged tiles included)
    uint32_t tiles_x, tiles_y, n_slots;
    int q_in;                          // shade reads queue[q_in], writes queue[q_in ^ 1]; extend walks queue[q_in ^ 1]
};

#define WF_NO_PATH 0xFFFFFFFFu

// per-hit data only (hit_record reads spheres / quads / meta / normals / materials); the extend kernel stages the
// traversal arrays with the megakernel's stage_scene<SMEM, KIND>
template <bool SMEM>
__device__ __forceinline__ DevScene wf_stage_scene(const DevScene& G) {
    if (!SMEM) return G;
    extern __shared__ float4 wf_smem_f4[];
    DevScene S = G;
    uint32_t off = 0;
    auto copy = [&](const float4* src, uint32_t n) -> const float4* {
        float4* dst = wf_smem_f4 + off;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
        off += n;
        return dst;
    };
    S.spheres = copy(G.spheres, G.n_spheres);                      // the shade kernel never touches the node arrays
    S.quads = copy(G.quads, 4 * G.n_quads);
    S.sphere_meta = (const uint4*)copy((const float4*)G.sphere_meta, G.n_spheres);
    S.quad_meta = (const uint4*)copy((const float4*)G.quad_meta, G.n_quads);
    S.quad_nn = copy(G.quad_nn, G.n_quads);
    S.materials = copy(G.materials, 2 * G.n_materials);
    __syncthreads();
    return S;
}

// path index -> pixel / sample; false for the padding paths of ragged edge tiles
__device__ __forceinline__ bool wf_decode_path(const WfArgs& A, unsigned long long p, uint32_t* x, uint32_t* y, uint32_t* s) {
    uint32_t lane = (uint32_t)(p & 31ull);
    unsigned long long q = p >> 5;
    uint32_t spp = A.p.spp;
    unsigned long long tile = q / spp;
    *s = (uint32_t)(q - tile * spp);
    uint32_t ty = (uint32_t)(tile / A.tiles_x), tx = (uint32_t)(tile - (unsigned long long)ty * A.tiles_x);
    *x = tx * 8 + (lane & 7u); *y = ty * 4 + (lane >> 3);
    return *x < A.p.width && *y < A.p.height;
}

// ------------------------------------------------------------------ shade + regenerate + compaction
template <bool SMEM>
__global__ void __launch_bounds__(WF_THREADS) k_wf_shade(const WfArgs A) {
    const DevScene S = wf_stage_scene<SMEM>(A.scene);
    const DevParams& P = A.p;
    WfCtl* ctl = A.ctl;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_in = ctl->q_count[A.q_in];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ctl->q_head = 0;                                   // for the extend kernel that follows
    if ((i & ~31u) >= n_in) return;                                // whole warp beyond the queue
    const bool have = i < n_in;
    uint32_t slot = have ? A.queue[A.q_in][i] : 0u;

    V3 o = mk(0, 0, 0), d = mk(0, 0, 1), atten = mk(1, 1, 1), color = mk(0, 0, 0);
    uint32_t pixel = 0, sample = 0, remain = 0, k = 0;
    bool alive = false, scatter = false;
    Surface sf; sf.kind = TRT_MAT_LIGHT; sf.front = false; sf.param = 0.f; sf.point = sf.normal = sf.rgb = mk(0, 0, 0);
    uint32_t n_paths = 0;
    if (have) {
        uint32_t rk = A.rk[slot];
        if (rk != WF_NO_PATH) {
            float4 a0 = A.a0[slot], a1 = A.a1[slot], a2 = A.a2[slot], a3 = A.a3[slot];
            o = mk(a0.x, a0.y, a0.z); d = mk(a0.w, a1.x, a1.y);
            Hit h; h.t = a1.z; h.ref = __float_as_uint(a1.w);
            atten = mk(a2.x, a2.y, a2.z); pixel = __float_as_uint(a2.w);
            color = mk(a3.x, a3.y, a3.z); sample = __float_as_uint(a3.w);
            remain = rk & 0xFFFFu; k = rk >> 16;
            // ---- cpu.rs:49-61 up to the scatter decision
            if (h.ref == TRT_REF_INVALID) color = color + atten * mk(P.bg[0], P.bg[1], P.bg[2]);
            else {
                sf = hit_record<SMEM>(S, h, o, d);
                V3 emission = (sf.kind == TRT_MAT_LIGHT) ? sf.rgb : mk(0.0f, 0.0f, 0.0f);
                color = color + atten * emission;
                scatter = sf.kind != TRT_MAT_LIGHT && remain > 1;
            }
            if (!scatter) {                                        // path finished: imager.rs:50 (f64 sums)
                double* dst = A.sum64 + 3 * (size_t)pixel;
                atomicAdd(dst, (double)color.x); atomicAdd(dst + 1, (double)color.y); atomicAdd(dst + 2, (double)color.z);
                ++n_paths;
            } else alive = true;
        }
    }
    // ---- refill finished / fresh slots with the next camera paths (warp-aggregated fetch, retry on padding paths)
    bool want = have && !alive;
    uint32_t px = 0, py = 0;
    bool regen = false;
    for (;;) {
        unsigned m = __ballot_sync(0xFFFFFFFFu, want);
        if (!m) break;
        int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(&ctl->next_path, (unsigned long long)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (want) {
            unsigned long long p = base + __popc(m & ((1u << lane) - 1u));
            if (p >= A.total_paths) want = false;                  // job exhausted: the slot dies
            else {
                uint32_t s;
                if (wf_decode_path(A, p, &px, &py, &s)) {
                    if (P.max_bounces == 0) ++n_paths;             // no ray is ever cast (cpu.rs:47): fetch the next one
                    else { pixel = py * P.width + px; sample = P.sample_offset + s; regen = true; want = false; }
                }
            }
        }
    }
    // ---- next ray: scatter (material/*.rs) or camera (pointgen.rs:37-52); one Philox call for both
    if (scatter || regen) {
        if (scatter) ++k; else k = 0;
        U4 r = philox4x32_10(pixel, sample, k, 0u, P.seed_lo, P.seed_hi);
        V3 dir;
        if (scatter) { dir = scatter_direction(sf, d, r); atten = atten * sf.rgb; o = sf.point; --remain; }
        else { camera_ray(A.cam, P, px, py, pixel, sample, r, &o, &dir); atten = mk(1, 1, 1); color = mk(0, 0, 0); remain = P.max_bounces; alive = true; }
        d = normalized(dir);
        A.a0[slot] = make_float4(o.x, o.y, o.z, d.x);
        A.a1[slot] = make_float4(d.y, d.z, TRT_INF, __uint_as_float(TRT_REF_INVALID));
        A.a2[slot] = make_float4(atten.x, atten.y, atten.z, __uint_as_float(pixel));
        A.a3[slot] = make_float4(color.x, color.y, color.z, __uint_as_float(sample));
        A.rk[slot] = (remain & 0xFFFFu) | (k << 16);
    } else if (have) A.rk[slot] = WF_NO_PATH;
    // ---- compaction of the slots that carry a ray: ballot + popc, one atomic per warp
    unsigned live = __ballot_sync(0xFFFFFFFFu, alive);
    if (live) {
        int leader = __ffs(live) - 1;
        unsigned base = 0;
        if ((int)lane == leader) base = atomicAdd(&ctl->q_count[A.q_in ^ 1], (unsigned)__popc(live));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (alive) A.queue[A.q_in ^ 1][base + __popc(live & ((1u << lane) - 1u))] = slot;
    }
    unsigned long long np = n_paths;
    for (int off = 16; off; off >>= 1) np += __shfl_xor_sync(0xFFFFFFFFu, np, off);
    if (lane == 0 && np) atomicAdd(&ctl->paths, np);
}

// ------------------------------------------------------------------ extend: persistent warps, per-lane ray replacement
// KIND as in the megakernel: 0 / 1 = TRT_BVH_REFERENCE / TRT_BVH_SAH on binary nodes, 4 / 3 = the same trees on their
// 4-wide nodes. The ordered kinds replay an order-dependent ray on the reference tree exactly as the megakernel does
// (DESIGN.md §4.3): inside the loop on wide nodes (begin_replay), by one ordered_replay call on binary nodes — so both
// variants return the same hit for every ray, and the images differ only in the order of the f64 accumulation.
template <int KIND, bool SMEM>
__global__ void __launch_bounds__(WF_THREADS, 3) k_wf_extend(const WfArgs A) {
    const DevScene S = stage_scene<SMEM, KIND>(A.scene);
    WfCtl* ctl = A.ctl;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t* queue = A.queue[A.q_in ^ 1];
    const uint32_t n = ctl->q_count[A.q_in ^ 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->q_count[A.q_in] = 0;   // the next shade pass appends to this queue
    Trav T; T.cur = TRT_REF_INVALID; T.sp = 0;
    TravStack K;
    uint32_t slot = 0; bool active = false, exhausted = false;
    uint32_t n_rays = 0;
    EventCounters ec;
    for (;;) {
        unsigned idle = __ballot_sync(0xFFFFFFFFu, !active);
        if (!exhausted && __popc(idle) >= WF_REFILL_MIN) {
            int leader = __ffs(idle) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(&ctl->q_head, (unsigned)__popc(idle));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active) {
                uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    slot = queue[idx];
                    float4 a0 = A.a0[slot], a1 = A.a1[slot];
                    trav_begin<false>(T, S, mk(a0.x, a0.y, a0.z), mk(a0.w, a1.x, a1.y), A.p.t_min, &ec);
                    active = true; ++n_rays;
                    if (T.cur == TRT_REF_INVALID) {                 // empty scene: an immediate miss
                        A.a1[slot] = make_float4(a1.x, a1.y, TRT_INF, __uint_as_float(TRT_REF_INVALID));
                        active = false;
                    }
                }
            }
            if (base + (unsigned)__popc(idle) >= n) exhausted = true;
            idle = __ballot_sync(0xFFFFFFFFu, !active);
        }
        if (idle == 0xFFFFFFFFu) { if (exhausted) break; else continue; }
        if (active) {
            if (KIND == 3 || KIND == 4) {
                trav4_any_step<KIND == 3, SMEM, false>(T, S, A.p.t_min, K, nullptr);
                if (KIND == 3 && T.cur == TRT_REF_INVALID && needs_replay(T, S)) begin_replay(T, S);     // rare: DESIGN.md §4.3
            } else trav_step<KIND == 1, SMEM, false>(T, S, K, &ec);
            if (T.cur == TRT_REF_INVALID) {
                Hit h = T.best;
                if (KIND == 1 && needs_replay(T, S)) h = ordered_replay(S, T.o, T.d, A.p.t_min);          // rare, binary walk
                A.a1[slot] = make_float4(T.d.y, T.d.z, h.t, __uint_as_float(h.ref));
                active = false;
            }
        }
    }
    unsigned long long r = n_rays;
    for (int off = 16; off; off >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, off);
    if (lane == 0 && r) atomicAdd(&ctl->rays, r);
}

__global__ void k_wf_init(uint32_t* queue0, uint32_t* rk, uint32_t n_slots, WfCtl* ctl, uint32_t n_start) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) { queue0[i] = i; rk[i] = WF_NO_PATH; }
    if (i == 0) { ctl->q_count[0] = n_start; ctl->q_count[1] = 0; ctl->q_head = 0; ctl->pad = 0; ctl->next_path = 0; ctl->rays = 0; ctl->paths = 0; }
}

// ------------------------------------------------------------------ launcher
namespace trtk {

size_t wavefront_scratch_bytes(uint32_t n_slots) {
    return (size_t)n_slots * (4 * 16 + 4 + 2 * 4) + 256;
}

cudaError_t launch_render_wavefront(const SceneView& v, const trt_camera& cam, const trt_params& p, double* d_sum64,
                                    void* d_scratch, uint32_t n_slots, void* d_ctl, int sm_count, cudaStream_t st,
                                    unsigned long long* h_rays_paths, uint32_t* launches_out) {
    // a TRT_BVH_FLAT scene also carries its reference-topology tree (trt_scene.cpp): the wavefront walks that one
    WfArgs A;
    A.scene = dev_scene(v); A.cam = dev_camera(cam); A.p = dev_params(p);
    char* b = (char*)d_scratch;
    A.a0 = (float4*)b; b += 16ull * n_slots;
    A.a1 = (float4*)b; b += 16ull * n_slots;
    A.a2 = (float4*)b; b += 16ull * n_slots;
    A.a3 = (float4*)b; b += 16ull * n_slots;
    A.rk = (uint32_t*)b; b += 4ull * n_slots;
    A.queue[0] = (uint32_t*)b; b += 4ull * n_slots;
    A.queue[1] = (uint32_t*)b;
    A.ctl = (WfCtl*)d_ctl; A.sum64 = d_sum64;
    A.tiles_x = (p.width + 7) / 8; A.tiles_y = (p.height + 3) / 4; A.n_slots = n_slots;
    A.total_paths = (unsigned long long)A.tiles_x * A.tiles_y * 32ull * p.spp;
    uint32_t launches = 0;
    if (A.total_paths == 0) { h_rays_paths[0] = h_rays_paths[1] = 0; *launches_out = 0; return cudaSuccess; }
    uint32_t n_start = (uint32_t)(A.total_paths < n_slots ? A.total_paths : n_slots);
    n_start = (n_start + 31u) & ~31u; if (n_start > n_slots) n_start = n_slots;
    k_wf_init<<<(n_slots + 255) / 256, 256, 0, st>>>(A.queue[0], A.rk, n_slots, A.ctl, n_start); ++launches;
    // shade stages the per-hit data, extend the traversal arrays of the kind it walks (wide nodes when the scene has them)
    size_t smem_shade = 16ull * (v.n_spheres + 4ull * v.n_quads + v.n_spheres + 2ull * v.n_quads + 2ull * v.n_materials);
    const bool shade_smem = smem_shade <= 64 * 1024;
    if (!shade_smem) smem_shade = 0;
    uint32_t kind = walk_kind(v, (p.flags & TRT_FLAG_BINARY_NODES) != 0);
    if (kind == 2) kind = 0;                                   // a flat-list scene is walked on its reference-topology tree
    size_t smem = scene_smem_bytes(v, kind);
    const bool use_smem = smem <= 64 * 1024;
    if (!use_smem) smem = 0;
    const void* f_shade = shade_smem ? (const void*)k_wf_shade<true> : (const void*)k_wf_shade<false>;
    const void* f_ext = nullptr;
#define WF_PICK(K) (use_smem ? (const void*)k_wf_extend<K, true> : (const void*)k_wf_extend<K, false>)
    f_ext = kind == 3 ? WF_PICK(3) : kind == 4 ? WF_PICK(4) : kind == 1 ? WF_PICK(1) : WF_PICK(0);
#undef WF_PICK
    cudaError_t e;
    if (smem_shade > 48 * 1024 && (e = cudaFuncSetAttribute(f_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_shade)) != cudaSuccess) return e;
    if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(f_ext, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    int bps = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, f_ext, WF_THREADS, smem)) != cudaSuccess) return e;
    if (bps < 1) bps = 1;
    int ext_grid = sm_count * bps;
    int shade_grid = (int)((n_slots + WF_THREADS - 1) / WF_THREADS);
    // host loop: one (shade, extend) pair per bounce of the pool; the control block is read back every CHECK pairs
    unsigned int* h_ctl = nullptr;
    if ((e = cudaMallocHost((void**)&h_ctl, sizeof(WfCtl))) != cudaSuccess) return e;
    const int CHECK = 8;
    A.q_in = 0;
    for (;;) {
        for (int it = 0; it < CHECK; ++it) {
            void* args[] = {(void*)&A};
            if ((e = cudaLaunchKernel(f_shade, dim3(shade_grid), dim3(WF_THREADS), args, smem_shade, st)) != cudaSuccess) break;
            if ((e = cudaLaunchKernel(f_ext, dim3(ext_grid), dim3(WF_THREADS), args, smem, st)) != cudaSuccess) break;
            launches += 2;
            A.q_in ^= 1;
        }
        if (e != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(h_ctl, A.ctl, sizeof(WfCtl), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        const WfCtl* c = (const WfCtl*)h_ctl;
        if (c->q_count[A.q_in] == 0 && c->next_path >= A.total_paths) { h_rays_paths[0] = c->rays; h_rays_paths[1] = c->paths; break; }
    }
    cudaFreeHost(h_ctl);
    *launches_out = launches;
    return e;
}

} // namespace trtk
