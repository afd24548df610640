// trt_capi.cpp — the C ABI of include/trt.h over the host scene (trt_scene.cpp) and the sm_100a kernels
// (trt_kernels.cu). No CPU fallback: every compute entry point needs a CUDA device and fails loudly without one.
#include "../../include/trt.h"
#include "trt_kernels.h"
#include "trt_scene.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <new>
#include <string>
#include <vector>

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
static int fail_cuda(const char* what, cudaError_t e) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return e == cudaErrorMemoryAllocation ? TRT_ERR_OOM : TRT_ERR_CUDA;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_cuda(#call, e_); } while (0)

struct DeviceCopy {
    uint64_t version = 0; void* buf = nullptr; size_t bytes = 0; trtk::SceneView view{}; uint32_t root_ref4 = 0; bool wide_ok = false;
    void* buf8 = nullptr; size_t bytes8 = 0;      // compressed 8-wide nodes + quad boxes (scenes whose wide nodes do not fit shared memory)
};
struct trt_scene {
    trt::HostScene host;
    std::map<int, DeviceCopy> on_device;   // device ordinal -> copy
    std::mutex host_mutex;                 // trt_scene_upload writes a little host-side state (device-computed leaf ranks, build
                                           // times, tree depth): trt_multi_render uploads to all devices at once (entries of
                                           // on_device are created beforehand on the calling thread; the map itself is not resized)
};
struct trt_ctx {
    int device = 0; int sm_count = 0; cudaStream_t stream = nullptr; cudaStream_t own_stream = nullptr; cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long* d_counters = nullptr;   // 8 x u64 (5 used) followed by the u32 work counter
    unsigned int* d_next_item = nullptr;
    // scratch reused across calls (grown on demand)
    // slots: 0 f64 sums, 1 f32 image, 2 8-bit image (trt_render), 3 wavefront path pool, 4..7 temporaries of trt_scene_upload's device
    // builds (primitive boxes, LBVH, reference order, wide-8 collapse) — kept, because cudaMalloc / cudaFree of ~200 MB per upload
    // cost 30...400 ms of wall time on a 10^6-primitive scene (scripts/e2e_probe.py) against 5 ms for the builds themselves
    void* scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; size_t scratch_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void* d_prim = nullptr; size_t prim_bytes = 0;   // records of the coherent first-bounce pass (one sample batch)
    char name[256];
};

static int ensure_scratch(trt_ctx* c, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (c->scratch_bytes[slot] < bytes) {
        if (c->scratch[slot]) { cudaFree(c->scratch[slot]); c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0; }
        CU(cudaMalloc(&c->scratch[slot], bytes));
        c->scratch_bytes[slot] = bytes;
    }
    *out = c->scratch[slot];
    return TRT_OK;
}

// the same for code that carries a cudaError_t along (trt_scene_upload's device builds)
static cudaError_t scratch_cu(trt_ctx* c, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (c->scratch_bytes[slot] < bytes) {
        if (c->scratch[slot]) { cudaFree(c->scratch[slot]); c->scratch[slot] = nullptr; c->scratch_bytes[slot] = 0; }
        cudaError_t e = cudaMalloc(&c->scratch[slot], bytes);
        if (e != cudaSuccess) { c->scratch[slot] = nullptr; return e; }
        c->scratch_bytes[slot] = bytes;
    }
    *out = c->scratch[slot];
    return cudaSuccess;
}

extern "C" {

const char* trt_last_error(void) { return g_last_error.c_str(); }
uint32_t trt_abi_version(void) { return TRT_ABI_VERSION; }

// ------------------------------------------------------------------ context
int trt_ctx_create(int device_ordinal, trt_ctx** out) {
    if (!out) return fail(TRT_ERR_INVALID, "trt_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(TRT_ERR_CUDA, std::string("trt_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device_ordinal < 0 || device_ordinal >= n) return fail(TRT_ERR_INVALID, "trt_ctx_create: device ordinal out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major != 10) return fail(TRT_ERR_UNSUPPORTED, std::string("trt_ctx_create: built for sm_100a only, device is ") + prop.name);
    CU(cudaSetDevice(device_ordinal));
    trt_ctx* c = new (std::nothrow) trt_ctx();
    if (!c) return fail(TRT_ERR_OOM, "trt_ctx_create: host allocation failed");
    c->device = device_ordinal; c->sm_count = prop.multiProcessorCount;
    std::snprintf(c->name, sizeof c->name, "%s", prop.name);
    cudaError_t ce = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    c->stream = c->own_stream;
    for (int i = 0; i < 4 && ce == cudaSuccess; ++i) ce = cudaEventCreate(&c->ev[i]);
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_counters, 8 * sizeof(unsigned long long) + 64);
    if (ce != cudaSuccess) { trt_ctx_destroy(c); return fail_cuda("trt_ctx_create", ce); }    // frees whatever was created
    c->d_next_item = (unsigned int*)(c->d_counters + 8);
    *out = c;
    return TRT_OK;
}
int trt_ctx_destroy(trt_ctx* c) {
    if (!c) return TRT_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 8; ++i) if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->d_prim) cudaFree(c->d_prim);
    if (c->d_counters) cudaFree(c->d_counters);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return TRT_OK;
}
int trt_ctx_device_name(trt_ctx* c, char* buf, size_t n) {
    if (!c || !buf || n == 0) return fail(TRT_ERR_INVALID, "trt_ctx_device_name: bad argument");
    std::snprintf(buf, n, "%s", c->name); return TRT_OK;
}
int trt_ctx_sm_count(trt_ctx* c, int* out) { if (!c || !out) return fail(TRT_ERR_INVALID, "trt_ctx_sm_count: bad argument"); *out = c->sm_count; return TRT_OK; }
int trt_ctx_synchronize(trt_ctx* c) { if (!c) return fail(TRT_ERR_INVALID, "trt_ctx_synchronize: null ctx"); CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->stream)); return TRT_OK; }

int trt_ctx_set_stream(trt_ctx* c, void* cuda_stream) {
    if (!c) return fail(TRT_ERR_INVALID, "trt_ctx_set_stream: null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream_t)cuda_stream;
    return TRT_OK;
}

// ------------------------------------------------------------------ scene (host only)
int trt_scene_create(trt_scene** out) {
    if (!out) return fail(TRT_ERR_INVALID, "trt_scene_create: out is null");
    *out = new (std::nothrow) trt_scene();
    return *out ? TRT_OK : fail(TRT_ERR_OOM, "trt_scene_create: host allocation failed");
}
int trt_scene_destroy(trt_scene* s) {
    if (!s) return TRT_OK;
    for (auto& kv : s->on_device) if (kv.second.buf) { cudaSetDevice(kv.first); cudaFree(kv.second.buf); if (kv.second.buf8) cudaFree(kv.second.buf8); }
    delete s; return TRT_OK;
}
static int add_mat(trt_scene* s, uint32_t kind, const float rgb[3], float param, uint32_t* id, const char* who) {
    if (!s || !rgb) return fail(TRT_ERR_INVALID, std::string(who) + ": null argument");
    uint32_t m = s->host.add_material(kind, rgb, param);
    if (id) *id = m;
    return TRT_OK;
}
int trt_add_lambertian(trt_scene* s, const float albedo[3], uint32_t* id) { return add_mat(s, TRT_MAT_LAMBERTIAN, albedo, 0.f, id, "trt_add_lambertian"); }
int trt_add_metal(trt_scene* s, const float albedo[3], float fuzz, uint32_t* id) { return add_mat(s, TRT_MAT_METAL, albedo, fuzz, id, "trt_add_metal"); }
int trt_add_dielectric(trt_scene* s, const float albedo[3], float ri, uint32_t* id) { return add_mat(s, TRT_MAT_DIELECTRIC, albedo, ri, id, "trt_add_dielectric"); }
int trt_add_light(trt_scene* s, const float color[3], uint32_t* id) { return add_mat(s, TRT_MAT_LIGHT, color, 0.f, id, "trt_add_light"); }

int trt_add_sphere(trt_scene* s, const float c[3], float r, uint32_t mat, uint32_t* id) {
    if (!s || !c) return fail(TRT_ERR_INVALID, "trt_add_sphere: null argument");
    std::string err; if (!s->host.add_sphere(c, r, mat, id, &err)) return fail(TRT_ERR_INVALID, err);
    return TRT_OK;
}
int trt_add_quad(trt_scene* s, const float corner[3], const float u[3], const float v[3], uint32_t mat, uint32_t* id) {
    if (!s || !corner || !u || !v) return fail(TRT_ERR_INVALID, "trt_add_quad: null argument");
    std::string err; if (!s->host.add_quad(corner, u, v, mat, id, &err)) return fail(TRT_ERR_INVALID, err);
    return TRT_OK;
}
int trt_add_prims(trt_scene* s, uint32_t n, const uint32_t* type, const float* data12, const uint32_t* mat) {
    if (!s || (n && (!type || !data12 || !mat))) return fail(TRT_ERR_INVALID, "trt_add_prims: null argument");
    s->host.prims.reserve(s->host.prims.size() + n);
    std::string err;
    for (uint32_t i = 0; i < n; ++i) {
        const float* d = data12 + 12 * (size_t)i;
        bool ok = type[i] == 0 ? s->host.add_sphere(d, d[3], mat[i], nullptr, &err)
                : type[i] == 1 ? s->host.add_quad(d, d + 3, d + 6, mat[i], nullptr, &err)
                : (err = "trt_add_prims: unknown primitive type", false);
        if (!ok) return fail(TRT_ERR_INVALID, err);
    }
    return TRT_OK;
}
int trt_scene_num_prims(trt_scene* s, uint32_t* out) { if (!s || !out) return fail(TRT_ERR_INVALID, "trt_scene_num_prims: null argument"); *out = (uint32_t)s->host.prims.size(); return TRT_OK; }

int trt_scene_commit(trt_scene* s, int bvh_kind) {
    if (!s) return fail(TRT_ERR_INVALID, "trt_scene_commit: null scene");
    std::string err;
    auto t0 = std::chrono::steady_clock::now();
    if (!s->host.commit(bvh_kind, &err)) return fail(bvh_kind >= 0 && bvh_kind <= 4 ? TRT_ERR_UNSUPPORTED : TRT_ERR_INVALID, err);
    s->host.host_commit_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    s->host.device_build_ms = 0.f;
    return TRT_OK;
}
int trt_scene_leaf_ranks(trt_scene* s, uint32_t* rank) {
    if (!s || !rank) return fail(TRT_ERR_INVALID, "trt_scene_leaf_ranks: null argument");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, "trt_scene_leaf_ranks: scene not committed");
    std::memcpy(rank, s->host.leaf_rank.data(), s->host.leaf_rank.size() * sizeof(uint32_t));
    return TRT_OK;
}
int trt_scene_build_times(trt_scene* s, float* host_commit_ms, float* device_build_ms) {
    if (!s) return fail(TRT_ERR_INVALID, "trt_scene_build_times: null scene");
    if (host_commit_ms) *host_commit_ms = s->host.host_commit_ms;
    if (device_build_ms) *device_build_ms = s->host.device_build_ms;
    return TRT_OK;
}
int trt_scene_bvh_info(trt_scene* s, uint32_t* n_nodes, uint32_t* max_depth, float* sah_cost) {
    if (!s) return fail(TRT_ERR_INVALID, "trt_scene_bvh_info: null scene");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, "trt_scene_bvh_info: scene not committed");
    if (n_nodes) *n_nodes = s->host.flat.n_nodes;
    if (max_depth) *max_depth = s->host.flat.max_depth;
    if (sah_cost) *sah_cost = s->host.flat.sah_cost;
    return TRT_OK;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int trt_scene_upload(trt_ctx* c, trt_scene* s) {
    if (!c || !s) return fail(TRT_ERR_INVALID, "trt_scene_upload: null argument");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, "trt_scene_upload: scene not committed (call trt_scene_commit)");
    CU(cudaSetDevice(c->device));
    DeviceCopy& dc = s->on_device[c->device];
    if (dc.buf && dc.version == s->host.version) return TRT_OK;
    const trt::FlatScene& f = s->host.flat;
    const bool device_build = f.bvh_kind == TRT_BVH_LBVH_DEVICE;
    const size_t n_prims = s->host.prims.size();
    size_t sz[11] = {device_build ? 64 * (n_prims - 1) : f.nodes.size() * 16, f.spheres.size() * 16, f.quads.size() * 16, f.sphere_meta.size() * 16, f.quad_meta.size() * 16, f.materials.size() * 16, f.flat.size() * 16, f.quad_nn.size() * 16, f.flat_cbox.size() * 16, (device_build || f.wide_on_device) ? 128 * (n_prims - 1) : f.nodes4.size() * 16, (device_build || f.ref_on_device) ? 128 * (n_prims - 1) : f.nodes_ref4.size() * 16};
    const void* src[11] = {f.nodes.data(), f.spheres.data(), f.quads.data(), f.sphere_meta.data(), f.quad_meta.data(), f.materials.data(), f.flat.data(), f.quad_nn.data(), f.flat_cbox.data(), f.nodes4.data(), f.nodes_ref4.data()};
    size_t off[11], total = 0;
    for (int i = 0; i < 11; ++i) { off[i] = total; total += align_up(sz[i] ? sz[i] : 16); }
    // a re-upload frees or overwrites memory that kernels of ANY stream of this device may still be reading (the streaming
    // sampler's two streams, a caller-owned stream): wait for all of them first. Rare (only after a re-commit).
    if (dc.buf) CU(cudaDeviceSynchronize());
    if (dc.buf && dc.bytes < total) { cudaFree(dc.buf); dc.buf = nullptr; }
    if (!dc.buf) { CU(cudaMalloc(&dc.buf, total)); dc.bytes = total; }
    for (int i = 0; i < 11; ++i) if (sz[i] && !(device_build && i == 0) && !((device_build || f.wide_on_device) && i == 9) && !((device_build || f.ref_on_device) && i == 10)) CU(cudaMemcpyAsync((char*)dc.buf + off[i], src[i], sz[i], cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    // Node::new (the reference's, level-parallel: trt_lbvh.cu k_ref_*) from the per-primitive boxes: ranks into meta.y (and back
    // to the host for trt_scene_leaf_ranks), the reference tree's wide nodes into the buffer
    auto ref_order = [&](const void* d_boxes, float* ms_out) -> cudaError_t {
        void* d_tmp2 = nullptr; uint32_t root_ref = 0;
        cudaError_t e = scratch_cu(c, 6, trtk::ref_order_tmp_bytes((uint32_t)n_prims), &d_tmp2);
        if (e == cudaSuccess) e = trtk::build_reference_order(d_boxes, (uint32_t)n_prims, (char*)dc.buf + off[3], (char*)dc.buf + off[4], (char*)dc.buf + off[10], d_tmp2, &root_ref, ms_out, c->stream);
        if (e != cudaSuccess) return e;
        dc.root_ref4 = root_ref;
        std::vector<trt_u4> meta(f.sphere_meta.size() + f.quad_meta.size());     // meta = (material, rank, prim_id, 0)
        if (sz[3]) e = cudaMemcpy(meta.data(), (char*)dc.buf + off[3], sz[3], cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && sz[4]) e = cudaMemcpy(meta.data() + f.sphere_meta.size(), (char*)dc.buf + off[4], sz[4], cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(s->host_mutex); for (const trt_u4& m : meta) if (m.z < s->host.leaf_rank.size()) s->host.leaf_rank[m.z] = m.y; }
        return e;
    };
    if (f.wide_on_device) {
        cudaError_t e = trtk::collapse_lbvh4((char*)dc.buf + off[0], (uint32_t)n_prims, (char*)dc.buf + off[9], false, c->stream);
        if (e != cudaSuccess) return fail_cuda("trt_scene_upload: wide nodes", e);
    }
    if (f.ref_on_device) {
        void* d_boxes = nullptr; float ms_ref = 0.f;
        cudaError_t e = scratch_cu(c, 4, f.lbvh_boxes.size() * 16, &d_boxes);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_boxes, f.lbvh_boxes.data(), f.lbvh_boxes.size() * 16, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = ref_order(d_boxes, &ms_ref);
        if (e != cudaSuccess) return fail_cuda("trt_scene_upload: device reference order", e);
        { std::lock_guard<std::mutex> lk(s->host_mutex); s->host.device_build_ms = ms_ref; }
    }
    if (device_build) {
        // the walked tree on the GPU: Morton codes + radix sort + Karras hierarchy + bottom-up boxes (trt_lbvh.cu)
        void *d_boxes = nullptr, *d_tmp = nullptr;
        cudaError_t e = scratch_cu(c, 4, f.lbvh_boxes.size() * 16, &d_boxes);   // pooled in the context: nothing to free on any path
        if (e == cudaSuccess) e = scratch_cu(c, 5, trtk::lbvh_tmp_bytes((uint32_t)n_prims), &d_tmp);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_boxes, f.lbvh_boxes.data(), f.lbvh_boxes.size() * 16, cudaMemcpyHostToDevice, c->stream);
        uint32_t depth = 0; float ms = 0.f;
        if (e == cudaSuccess) e = trtk::build_lbvh(d_boxes, (uint32_t)n_prims, (char*)dc.buf + off[0], (char*)dc.buf + off[9], d_tmp, &depth, &ms, c->stream);
        // the wide records: plain two-level collapse (opening the largest children instead, as the host does for the SAH tree,
        // measured no gain on the device-built tree: 10^6 spheres 2 100 vs 2 112 Mrays/s)
        const bool greedy4 = false;
        if (e == cudaSuccess) e = trtk::collapse_lbvh4((char*)dc.buf + off[0], (uint32_t)n_prims, (char*)dc.buf + off[9], greedy4, c->stream);
        dc.wide_ok = greedy4 || 3 * ((depth + 1) / 2) + 1 <= TRT_MAX_STACK;
        // Node::new itself on the device: the reference leaf order (tie rule) and the reference tree's wide nodes (replay)
        float ms_ref = 0.f;
        if (e == cudaSuccess) e = ref_order(d_boxes, &ms_ref);
        if (e != cudaSuccess) return fail_cuda("trt_scene_upload: device BVH build", e);
        { std::lock_guard<std::mutex> lk(s->host_mutex); s->host.flat.max_depth = depth; s->host.device_build_ms = ms + ms_ref; }
        if (depth > TRT_MAX_STACK) return fail(TRT_ERR_UNSUPPORTED, "trt_scene_upload: device-built tree deeper than the traversal stack; commit with TRT_BVH_SAH");
    }
    trtk::SceneView& v = dc.view;
    char* b = (char*)dc.buf;
    v.nodes = b + off[0]; v.spheres = b + off[1]; v.quads = b + off[2]; v.sphere_meta = b + off[3]; v.quad_meta = b + off[4]; v.quad_nn = b + off[7]; v.materials = b + off[5];
    v.root = f.root; v.n_nodes = f.n_nodes; v.n_spheres = (uint32_t)f.spheres.size(); v.n_quads = (uint32_t)f.quad_meta.size();
    v.n_materials = (uint32_t)(f.materials.size() / 2); v.kind = device_build ? 1u : (uint32_t)f.bvh_kind;   // the device-built tree is walked near-child-first like SAH
    v.flat = f.flat.empty() ? nullptr : b + off[6]; v.n_flat = (uint32_t)(f.flat.size() / TRT_FLAT_F4); v.flat_qmask[0] = f.flat_qmask[0]; v.flat_qmask[1] = f.flat_qmask[1];
    v.nodes4 = f.nodes4.empty() ? nullptr : b + off[9]; v.n_nodes4 = f.n_nodes4;
    v.nodes_ref4 = f.nodes_ref4.empty() ? nullptr : b + off[10]; v.n_nodes_ref4 = f.n_nodes_ref4; v.root_ref4 = 0;
    if (device_build || f.ref_on_device) { v.nodes_ref4 = b + off[10]; v.n_nodes_ref4 = (uint32_t)(n_prims - 1); v.root_ref4 = dc.root_ref4; }
    if ((device_build && dc.wide_ok) || f.wide_on_device) { v.nodes4 = b + off[9]; v.n_nodes4 = (uint32_t)(n_prims - 1); }
    v.flat_cbox = f.flat_cbox.empty() ? nullptr : b + off[8]; v.flat_cbox_host = f.flat_cbox.empty() ? nullptr : f.flat_cbox.data(); v.flat_bound = f.flat_bound;
    const float inf = std::numeric_limits<float>::infinity();
    for (int i = 0; i < 3; ++i) { v.root_min[i] = inf; v.root_max[i] = -inf; }
    for (const trt::HostPrim& p : s->host.prims)
        for (int i = 0; i < 3; ++i) { v.root_min[i] = fminf(v.root_min[i], p.bmin[i]); v.root_max[i] = fmaxf(v.root_max[i], p.bmax[i]); }
    // Compressed 8-wide nodes (trt_lbvh.cu k_cw8_level) for ordered trees that are walked in global memory: built on
    // the device from the binary records that are there anyway (uploaded SAH tree or device-built tree). OPT-IN (TRT_WIDE8=1 in the
    // environment when the scene is uploaded): bit-identical results, a third fewer node fetches and 43 % fewer L1 sectors, but
    // the decode makes the walk ALU-pipe-bound on B200 — 10^6 spheres 1 994 vs 2 317 Mrays/s on the 4-wide records
    // (profiles/r2g_tree_c4_2spp_wide8 / _wide4_ld256) — so the 4-wide walk stays the default.
    v.nodes8 = nullptr; v.n_nodes8 = 0; v.quad_box = nullptr; v.bound8 = -1.0f;
    {
        float bound = 0.f; bool finite = n_prims >= 2;
        for (int i = 0; i < 3; ++i) { bound = fmaxf(bound, fmaxf(fabsf(v.root_min[i]), fabsf(v.root_max[i]))); if (!(fabsf(v.root_min[i]) <= 1073741824.0f && fabsf(v.root_max[i]) <= 1073741824.0f)) finite = false; }
        const char* on8 = std::getenv("TRT_WIDE8");
        const bool want8 = finite && v.kind == 1 && v.nodes4 && v.nodes_ref4 && trtk::scene_smem_bytes(v, 3) > 64 * 1024 && on8 && on8[0] == '1';
        if (want8) {
            const size_t nq = v.n_quads, node_bytes = align_up(96 * (n_prims - 1)), qb_bytes = align_up(32 * (nq ? nq : 1));
            const size_t need = node_bytes + qb_bytes;
            if (dc.buf8 && dc.bytes8 < need) { cudaFree(dc.buf8); dc.buf8 = nullptr; dc.bytes8 = 0; }
            if (!dc.buf8) { CU(cudaMalloc(&dc.buf8, need)); dc.bytes8 = need; }
            std::vector<trt_f4> qb(2 * nq);
            for (const trt::HostPrim& p : s->host.prims) if (p.type == 1 && p.type_index < nq) {     // same slot as the quad's geometry (trt_scene.cpp)
                qb[2 * (size_t)p.type_index] = trt_f4{p.bmin[0], p.bmin[1], p.bmin[2], 0.f}; qb[2 * (size_t)p.type_index + 1] = trt_f4{p.bmax[0], p.bmax[1], p.bmax[2], 0.f};
            }
            void* d_tmp8 = nullptr; uint32_t n8 = 0, levels8 = 0;
            cudaError_t e = cudaSuccess;
            if (nq) e = cudaMemcpyAsync((char*)dc.buf8 + node_bytes, qb.data(), 32 * nq, cudaMemcpyHostToDevice, c->stream);
            if (e == cudaSuccess) e = scratch_cu(c, 7, trtk::cw8_tmp_bytes((uint32_t)n_prims), &d_tmp8);
            if (e == cudaSuccess) e = trtk::build_cw8(v.nodes, (uint32_t)n_prims, v.root, dc.buf8, d_tmp8, &n8, &levels8, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return fail_cuda("trt_scene_upload: compressed 8-wide nodes", e);
            v.nodes8 = dc.buf8; v.n_nodes8 = n8; v.quad_box = (char*)dc.buf8 + node_bytes; v.bound8 = bound;
        }
    }
    dc.version = s->host.version;
    return TRT_OK;
}

// ------------------------------------------------------------------ camera (host only): Camera::new, camera.rs:17-56
int trt_camera_make(float focus_distance, float defocus_angle, const float position[3], const float look_at[3],
                    const float up[3], float vertical_fov, uint32_t width, uint32_t height, trt_camera* out) {
    if (!position || !look_at || !up || !out) return fail(TRT_ERR_INVALID, "trt_camera_make: null argument");
    if (width == 0 || height == 0) return fail(TRT_ERR_INVALID, "trt_camera_make: empty image");
    const float rads_per_deg = 3.14159265358979323846f / 180.0f;             // f32::to_radians
    float viewport_height = 2.0f * focus_distance * std::tan((vertical_fov * rads_per_deg) / 2.0f);
    float aspect_ratio = (float)width / (float)height;
    float viewport_width = aspect_ratio * viewport_height;
    auto norm3 = [](const float v[3], float o[3]) { float l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); o[0] = v[0] / l; o[1] = v[1] / l; o[2] = v[2] / l; };
    auto cross3 = [](const float a[3], const float b[3], float o[3]) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    float pl[3] = {position[0] - look_at[0], position[1] - look_at[1], position[2] - look_at[2]};
    float w[3], u[3], v[3], t[3];
    norm3(pl, w); cross3(up, w, t); norm3(t, u); cross3(w, u, t); norm3(t, v);
    float defocus_radius = focus_distance * std::tan((defocus_angle * rads_per_deg) / 2.0f);
    for (int i = 0; i < 3; ++i) {
        out->position[i] = position[i];
        out->forward[i] = w[i] * focus_distance;
        out->horizontal[i] = u[i] * viewport_width;
        out->vertical[i] = v[i] * viewport_height;
        out->viewport_upper_left[i] = position[i] - out->horizontal[i] / 2.0f + out->vertical[i] / 2.0f - out->forward[i];
        out->defocus_disk_u[i] = u[i] * defocus_radius;
        out->defocus_disk_v[i] = v[i] * defocus_radius;
    }
    out->width = width; out->height = height;
    return TRT_OK;
}

// ------------------------------------------------------------------ rendering
static int check_render_args(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, const char* who) {
    if (!c || !s || !cam || !p) return fail(TRT_ERR_INVALID, std::string(who) + ": null argument");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, std::string(who) + ": scene not committed (call trt_scene_commit)");
    if (p->width != cam->width || p->height != cam->height) return fail(TRT_ERR_INVALID, std::string(who) + ": params and camera disagree on the image size");
    if (p->width == 0 || p->height == 0) return fail(TRT_ERR_INVALID, std::string(who) + ": empty image");
    if ((uint64_t)p->width * p->height > 0x7FFFFFFFull) return fail(TRT_ERR_INVALID, std::string(who) + ": image too large");
    if (p->variant != TRT_VARIANT_MEGAKERNEL && p->variant != TRT_VARIANT_WAVEFRONT) return fail(TRT_ERR_INVALID, std::string(who) + ": unknown variant");
    return TRT_OK;
}

static int render_into(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, double* d_sum64, trt_stats* stats);
static int render_into(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, double* d_sum64, trt_stats* stats) {
    int rc = trt_scene_upload(c, s); if (rc) return rc;
    const trtk::SceneView& v = s->on_device[c->device].view;
    if (p->variant == TRT_VARIANT_WAVEFRONT) {
        if (p->max_bounces > 0xFFFFu) return fail(TRT_ERR_UNSUPPORTED, "wavefront variant: max_bounces above 65535");
        uint32_t n_slots = 1u << 20;                       // 68 B per slot: 68 MB of path state, L2-resident on B200
        if (const char* e = std::getenv("TRT_WF_SLOTS")) { long x = std::atol(e); if (x >= 1024 && x <= (1l << 26)) n_slots = (uint32_t)x & ~31u; }
        void* d_state = nullptr; int rc2;
        if ((rc2 = ensure_scratch(c, 3, trtk::wavefront_scratch_bytes(n_slots), &d_state))) return rc2;
        unsigned long long rp[2] = {0, 0}; uint32_t launches = 0;
        CU(cudaEventRecord(c->ev[0], c->stream));
        CU(trtk::launch_render_wavefront(v, *cam, *p, d_sum64, d_state, n_slots, c->d_counters, c->sm_count, c->stream, rp, &launches));
        CU(cudaEventRecord(c->ev[1], c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (stats) {
            float ms = 0.f; CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
            stats->rays = rp[0]; stats->paths = rp[1]; stats->node_visits = stats->sphere_tests = stats->quad_tests = 0;
            stats->device_ms = ms; stats->launches = launches; stats->finalize_ms = 0.f; stats->reserved = 0;
        }
        return TRT_OK;
    }
    CU(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long) + 64, c->stream));
    // The coherent first-bounce pass (k_primary) hands the path loop one 32-byte record per path through HBM. Samples are
    // rendered in batches whose records fit the scratch budget (TRT_PRIMARY_BUDGET_MB, default 16 GB, at most half of the
    // free memory); every batch is one k_primary + one megakernel launch over its sample range. A pixel's f64 sum is continued
    // from batch to batch in sample order (trt_kernels.cu fetch_items), so batching does not change a single bit.
    // It is used where it was measured to pay (trtk::primary_pass_pays: wide-node trees staged in shared memory, final scene
    // +7 %; the flat list and the trees in L2 lose 1-14 % with it and keep inline primary rays) or when TRT_FLAG_PRIMARY_PASS asks.
    const bool want_prim = p->spp > 0 && p->max_bounces > 0 && !(p->flags & (TRT_FLAG_COUNT_EVENTS | TRT_FLAG_COUNT_WALKED | TRT_FLAG_INLINE_PRIMARY)) &&
                           ((p->flags & TRT_FLAG_PRIMARY_PASS) || trtk::primary_pass_pays(v, *p));
    uint32_t batch = p->spp; void* d_prim = nullptr;
    if (want_prim) {
        size_t budget = (size_t)16384 << 20;
        if (const char* e = std::getenv("TRT_PRIMARY_BUDGET_MB")) { long long mb = std::atoll(e); if (mb >= 0) budget = (size_t)mb << 20; }
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = std::min(budget, std::max(c->prim_bytes, free_b / 2 + c->prim_bytes / 2));
        const size_t per_sample = trtk::primary_record_bytes(*p);
        size_t k = per_sample ? budget / per_sample : 0;
        if (k >= 1) {
            batch = (uint32_t)std::min<size_t>(k, p->spp);
            const size_t need = per_sample * batch;
            if (c->prim_bytes < need) {
                if (c->d_prim) { cudaFree(c->d_prim); c->d_prim = nullptr; c->prim_bytes = 0; }
                if (cudaMalloc(&c->d_prim, need) == cudaSuccess) c->prim_bytes = need;
                else { cudaGetLastError(); c->d_prim = nullptr; }            // no room: render with inline primary rays instead
            }
            d_prim = c->d_prim;
        }
        if (!d_prim) batch = p->spp;
    }
    uint32_t launches = 0;
    CU(cudaEventRecord(c->ev[0], c->stream));
    for (uint32_t off = 0; off < p->spp || (off == 0 && p->spp == 0); off += batch) {
        trt_params q = *p;
        q.spp = std::min(batch, p->spp - off); q.sample_offset = p->sample_offset + off;
        if (off) CU(cudaMemsetAsync(c->d_next_item, 0, sizeof(unsigned int), c->stream));
        CU(trtk::launch_render_mega(v, *cam, q, d_sum64, c->d_counters, c->d_next_item, c->sm_count, c->stream, d_prim));
        launches += d_prim ? 2u : 1u;
        if (p->spp == 0) break;
    }
    CU(cudaEventRecord(c->ev[1], c->stream));
    if (stats) {
        unsigned long long h[8];
        CU(cudaMemcpyAsync(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0.f; CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        stats->rays = h[0]; stats->paths = h[1]; stats->node_visits = h[2]; stats->sphere_tests = h[3]; stats->quad_tests = h[4];
        stats->device_ms = ms; stats->launches = launches; stats->finalize_ms = 0.f; stats->reserved = 0;
    }
    return TRT_OK;
}

int trt_render_accumulate(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, double* d_sum64, trt_stats* stats) {
    int rc = check_render_args(c, s, cam, p, "trt_render_accumulate"); if (rc) return rc;
    if (!d_sum64) return fail(TRT_ERR_INVALID, "trt_render_accumulate: d_sum64 is null");
    CU(cudaSetDevice(c->device));
    return render_into(c, s, cam, p, d_sum64, stats);
}

int trt_finalize(trt_ctx* c, const double* d_sum64, uint32_t width, uint32_t height, uint32_t spp_total, float* d_linear, uint8_t* d_rgb8, trt_stats* stats) {
    if (!c || !d_sum64) return fail(TRT_ERR_INVALID, "trt_finalize: null argument");
    if (spp_total == 0) return fail(TRT_ERR_INVALID, "trt_finalize: spp_total is zero");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev[2], c->stream));
    CU(trtk::launch_finalize(d_sum64, (size_t)width * height * 3, spp_total, d_linear, d_rgb8, c->stream));
    CU(cudaEventRecord(c->ev[3], c->stream));
    if (stats) {
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0.f; CU(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
        stats->finalize_ms = ms; stats->launches += 1;
    }
    return TRT_OK;
}

int trt_render(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, float* out_linear, uint8_t* out_rgb8, trt_stats* stats) {
    int rc = check_render_args(c, s, cam, p, "trt_render"); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    size_t n = (size_t)p->width * p->height * 3;
    void *d_sum = nullptr, *d_lin = nullptr, *d_u8 = nullptr;
    if ((rc = ensure_scratch(c, 0, n * sizeof(double), &d_sum))) return rc;
    if ((rc = ensure_scratch(c, 1, n * sizeof(float), &d_lin))) return rc;
    if ((rc = ensure_scratch(c, 2, n, &d_u8))) return rc;
    CU(cudaMemsetAsync(d_sum, 0, n * sizeof(double), c->stream));
    trt_stats st; std::memset(&st, 0, sizeof st);
    if ((rc = render_into(c, s, cam, p, (double*)d_sum, &st))) return rc;
    uint32_t spp_total = p->spp_total ? p->spp_total : p->spp;
    if (spp_total == 0) spp_total = 1;
    if ((rc = trt_finalize(c, (const double*)d_sum, p->width, p->height, spp_total, out_linear ? (float*)d_lin : nullptr, out_rgb8 ? (uint8_t*)d_u8 : nullptr, &st))) return rc;
    if (out_linear) CU(cudaMemcpyAsync(out_linear, d_lin, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (out_rgb8) CU(cudaMemcpyAsync(out_rgb8, d_u8, n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (stats) *stats = st;
    return TRT_OK;
}

// ------------------------------------------------------------------ multi-GPU behind the boundary
// Renderer::render (renderer/renderer.rs:37-79) on the GPUs of one box, from ONE call of a C / C++ / Rust host: one host
// thread and one trt_ctx per device, samples-per-pixel partitioned contiguously (rank k renders global sample indices
// [offset + k*spp/N, ...), Philox keyed by the global index, so the union of paths is the 1-GPU set), every GPU keeps the f64
// sums of its samples, then the fused reduce-scatter + tone-map kernel of trt_multi.cu: GPU k pulls slice k of every peer's
// buffer over NVLink (peer-mapped loads), adds in rank order, normalises and tone-maps, and copies its slice of the result
// straight into the caller's host buffers — N PCIe links in parallel. No NCCL, no Python.
namespace {
struct Barrier {            // C++17 has no std::barrier
    std::mutex m; std::condition_variable cv; int n, waiting = 0; unsigned gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
struct MultiSlot {
    int device = 0; trt_ctx* ctx = nullptr;
    double* d_sum = nullptr; size_t cap_values = 0;              // whole-image f64 sums of this rank's samples
    float* d_lin = nullptr; uint8_t* d_u8 = nullptr; double* d_stage = nullptr; size_t cap_slice = 0, cap_stage = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rc = TRT_OK; std::string err; trt_stats st{};
};
}
struct trt_multi {
    std::vector<MultiSlot> slot;
    bool direct = true;          // every pair of distinct devices can peer-load: the reduce reads the peers' buffers in place
};

int trt_multi_destroy(trt_multi* m) {
    if (!m) return TRT_OK;
    for (MultiSlot& k : m->slot) {
        if (!k.ctx) continue;
        cudaSetDevice(k.device);
        cudaStreamSynchronize(k.ctx->stream);
        if (k.d_sum) cudaFree(k.d_sum);
        if (k.d_lin) cudaFree(k.d_lin);
        if (k.d_u8) cudaFree(k.d_u8);
        if (k.d_stage) cudaFree(k.d_stage);
        if (k.ev0) cudaEventDestroy(k.ev0);
        if (k.ev1) cudaEventDestroy(k.ev1);
        trt_ctx_destroy(k.ctx);
    }
    delete m;
    return TRT_OK;
}

int trt_multi_create(const int* devices, int n_devices, trt_multi** out) {
    if (!out) return fail(TRT_ERR_INVALID, "trt_multi_create: out is null");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > TRT_MULTI_MAX_DEVICES) return fail(TRT_ERR_INVALID, "trt_multi_create: 1..16 devices");
    trt_multi* m = new (std::nothrow) trt_multi();
    if (!m) return fail(TRT_ERR_OOM, "trt_multi_create: host allocation failed");
    m->slot.resize((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) {
        MultiSlot& k = m->slot[(size_t)i];
        k.device = devices[i];
        int rc = trt_ctx_create(devices[i], &k.ctx);
        if (rc) { std::string e = g_last_error; trt_multi_destroy(m); return fail(rc, e); }
        cudaError_t ce = cudaEventCreate(&k.ev0);
        if (ce == cudaSuccess) ce = cudaEventCreate(&k.ev1);
        if (ce != cudaSuccess) { trt_multi_destroy(m); return fail_cuda("trt_multi_create", ce); }
    }
    // peer mapping in both directions for every pair of distinct devices (NVLink / NVSwitch on a B200 box)
    for (int i = 0; i < n_devices && m->direct; ++i)
        for (int j = 0; j < n_devices; ++j) {
            const int a = devices[i], b = devices[j];
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { m->direct = false; break; }
            cudaSetDevice(a);
            cudaError_t ce = cudaDeviceEnablePeerAccess(b, 0);
            if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();      // fine: clear the sticky-less error
            else if (ce != cudaSuccess) { cudaGetLastError(); m->direct = false; break; }
        }
    if (const char* e = std::getenv("TRT_MULTI_STAGED")) if (e[0] == '1') m->direct = false;    // test hook: force the copy path
    *out = m;
    return TRT_OK;
}
int trt_multi_device_count(trt_multi* m, int* n_devices, int* peer_direct) {
    if (!m || !n_devices) return fail(TRT_ERR_INVALID, "trt_multi_device_count: null argument");
    *n_devices = (int)m->slot.size();
    if (peer_direct) *peer_direct = m->direct ? 1 : 0;
    return TRT_OK;
}

static int render_into(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, double* d_sum64, trt_stats* stats);

int trt_multi_render(trt_multi* m, trt_scene* s, const trt_camera* cam, const trt_params* p, float* out_linear, uint8_t* out_rgb8, trt_stats* stats) {
    if (!m || m->slot.empty()) return fail(TRT_ERR_INVALID, "trt_multi_render: null handle");
    int rc = check_render_args(m->slot[0].ctx, s, cam, p, "trt_multi_render"); if (rc) return rc;
    const int N = (int)m->slot.size();
    const size_t n_values = (size_t)p->width * p->height * 3, pairs = (n_values + 1) / 2;
    uint32_t spp_total = p->spp_total ? p->spp_total : p->spp;
    if (spp_total == 0) spp_total = 1;
    std::vector<size_t> cut((size_t)N + 1);                         // slice k = [cut[k], cut[k+1]), even boundaries (double2 loads)
    for (int k = 0; k <= N; ++k) cut[(size_t)k] = std::min(n_values, 2 * (pairs * (size_t)k / (size_t)N));
    cut[(size_t)N] = n_values;
    size_t max_slice = 0;
    for (int k = 0; k < N; ++k) max_slice = std::max(max_slice, cut[(size_t)k + 1] - cut[(size_t)k]);
    max_slice = (max_slice + 1) & ~(size_t)1;
    // buffers: serially, on the calling thread; the scene's device copies are made by the workers, all devices at once
    // (one upload per DEVICE: a device listed twice is uploaded by its first slot, the others wait at the barrier below)
    for (MultiSlot& k : m->slot) {
        CU(cudaSetDevice(k.device));
        if (k.cap_values < n_values) {
            if (k.d_sum) { cudaFree(k.d_sum); k.d_sum = nullptr; k.cap_values = 0; }
            CU(cudaMalloc((void**)&k.d_sum, (n_values + 2) * sizeof(double))); k.cap_values = n_values;
        }
        if (k.cap_slice < max_slice) {
            if (k.d_lin) { cudaFree(k.d_lin); k.d_lin = nullptr; } if (k.d_u8) { cudaFree(k.d_u8); k.d_u8 = nullptr; } k.cap_slice = 0;
            CU(cudaMalloc((void**)&k.d_lin, max_slice * sizeof(float))); CU(cudaMalloc((void**)&k.d_u8, max_slice)); k.cap_slice = max_slice;
        }
        if (!m->direct && k.cap_stage < max_slice * (size_t)N) {
            if (k.d_stage) { cudaFree(k.d_stage); k.d_stage = nullptr; k.cap_stage = 0; }
            CU(cudaMalloc((void**)&k.d_stage, max_slice * (size_t)N * sizeof(double))); k.cap_stage = max_slice * (size_t)N;
        }
        (void)s->on_device[k.device];                                // create the map entry here: the workers only look it up
    }
    Barrier bar(N);
    std::vector<int> failed_flag(1, 0);
    std::mutex fail_m;
    auto worker = [&](int r) {
        MultiSlot& k = m->slot[(size_t)r];
        std::memset(&k.st, 0, sizeof k.st); k.rc = TRT_OK; k.err.clear();
        auto bad = [&](int code) { k.rc = code; k.err = g_last_error; std::lock_guard<std::mutex> lk(fail_m); failed_flag[0] = 1; };
        auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && k.rc == TRT_OK) bad(fail_cuda(what, e)); return e == cudaSuccess; };
        cudaStream_t st = k.ctx->stream;
        // ---- the scene on this device (upload + device-side builds), once per device
        bool first_of_device = true;
        for (int q = 0; q < r; ++q) if (m->slot[(size_t)q].device == k.device) first_of_device = false;
        if (first_of_device && cu(cudaSetDevice(k.device), "cudaSetDevice")) { int rc2 = trt_scene_upload(k.ctx, s); if (rc2) bad(rc2); }
        bar.wait();
        if (failed_flag[0]) return;
        // ---- this rank's share of the samples: contiguous, balanced (first spp % N ranks get one more)
        const uint32_t base = p->spp / (uint32_t)N, extra = p->spp % (uint32_t)N;
        const uint32_t cnt = base + ((uint32_t)r < extra ? 1u : 0u), off = (uint32_t)r * base + std::min((uint32_t)r, extra);
        if (cu(cudaSetDevice(k.device), "cudaSetDevice") && cu(cudaMemsetAsync(k.d_sum, 0, n_values * sizeof(double), st), "cudaMemsetAsync")) {
            if (cnt) {
                trt_params q = *p; q.spp = cnt; q.sample_offset = p->sample_offset + off; q.spp_total = spp_total;
                int rc2 = render_into(k.ctx, s, cam, &q, k.d_sum, &k.st);       // synchronises the stream (stats)
                if (rc2) bad(rc2);
            } else cu(cudaStreamSynchronize(st), "cudaStreamSynchronize");
        }
        bar.wait();                                                  // every rank's sums are complete and visible
        if (failed_flag[0]) return;
        // ---- reduce-scatter + epilogue of slice r, then the slice goes home over this GPU's own PCIe link
        const size_t b = cut[(size_t)r], e = cut[(size_t)r + 1], len = e - b;
        const double* src[TRT_MULTI_MAX_DEVICES];
        for (int q = 0; q < N; ++q) {
            const MultiSlot& o = m->slot[(size_t)q];
            if (m->direct || o.device == k.device) src[q] = o.d_sum;
            else {      // no peer mapping: copy the peer's slice into a local staging area first
                double* dst = k.d_stage + (size_t)q * max_slice;
                if (len) cu(cudaMemcpyPeerAsync(dst, k.device, o.d_sum + b, o.device, len * sizeof(double), st), "cudaMemcpyPeerAsync");
                src[q] = dst - b;                                    // indexed with whole-image indices by the kernel
            }
        }
        cu(cudaEventRecord(k.ev0, st), "cudaEventRecord");
        cu(trtk::launch_reduce_scatter_finalize(src, N, b, e, spp_total, k.d_lin, k.d_u8, k.ctx->sm_count, st), "launch_reduce_scatter_finalize");
        cu(cudaEventRecord(k.ev1, st), "cudaEventRecord");
        if (out_linear && len) cu(cudaMemcpyAsync(out_linear + b, k.d_lin, len * sizeof(float), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
        if (out_rgb8 && len) cu(cudaMemcpyAsync(out_rgb8 + b, k.d_u8, len, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync");
        cu(cudaStreamSynchronize(st), "cudaStreamSynchronize");
        float ms = 0.f;
        if (k.rc == TRT_OK && cu(cudaEventElapsedTime(&ms, k.ev0, k.ev1), "cudaEventElapsedTime")) { k.st.finalize_ms = ms; k.st.launches += len ? 1 : 0; }
    };
    std::vector<std::thread> th;
    for (int r = 1; r < N; ++r) th.emplace_back(worker, r);
    worker(0);
    for (std::thread& t : th) t.join();
    trt_stats tot; std::memset(&tot, 0, sizeof tot);
    for (const MultiSlot& k : m->slot) {
        if (k.rc) return fail(k.rc, "trt_multi_render (device " + std::to_string(k.device) + "): " + k.err);
        tot.paths += k.st.paths; tot.rays += k.st.rays; tot.node_visits += k.st.node_visits; tot.sphere_tests += k.st.sphere_tests; tot.quad_tests += k.st.quad_tests;
        tot.launches += k.st.launches;
        tot.device_ms = std::max(tot.device_ms, k.st.device_ms);            // the ranks run side by side: the slowest one counts
        tot.finalize_ms = std::max(tot.finalize_ms, k.st.finalize_ms);
    }
    if (stats) *stats = tot;
    CU(cudaSetDevice(m->slot[0].device));
    return TRT_OK;
}

int trt_render_multi(const int* devices, int n_devices, trt_scene* s, const trt_camera* cam, const trt_params* p,
                     float* out_linear, uint8_t* out_rgb8, trt_stats* stats) {
    trt_multi* m = nullptr;
    int rc = trt_multi_create(devices, n_devices, &m); if (rc) return rc;
    rc = trt_multi_render(m, s, cam, p, out_linear, out_rgb8, stats);
    std::string e = g_last_error;
    trt_multi_destroy(m);
    if (rc) g_last_error = e;
    return rc;
}

int trt_render_aov(trt_ctx* c, trt_scene* s, const trt_camera* cam, const trt_params* p, uint32_t* prim_id, float* t, float* radiance, uint32_t* n_rays) {
    int rc = check_render_args(c, s, cam, p, "trt_render_aov"); if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if ((rc = trt_scene_upload(c, s))) return rc;
    const trtk::SceneView& v = s->on_device[c->device].view;
    size_t n = (size_t)p->width * p->height * p->spp;
    void *d_prim = nullptr, *d_t = nullptr, *d_rad = nullptr, *d_nr = nullptr;
    if (prim_id && (rc = ensure_scratch(c, 0, n * 4, &d_prim))) return rc;
    if (t && (rc = ensure_scratch(c, 1, n * 4, &d_t))) return rc;
    if (radiance && (rc = ensure_scratch(c, 2, n * 12, &d_rad))) return rc;
    if (n_rays && (rc = ensure_scratch(c, 3, n * 4, &d_nr))) return rc;
    CU(trtk::launch_aov(v, *cam, *p, (uint32_t*)d_prim, (float*)d_t, (float*)d_rad, (uint32_t*)d_nr, c->stream));
    if (prim_id) CU(cudaMemcpyAsync(prim_id, d_prim, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (t) CU(cudaMemcpyAsync(t, d_t, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (radiance) CU(cudaMemcpyAsync(radiance, d_rad, n * 12, cudaMemcpyDeviceToHost, c->stream));
    if (n_rays) CU(cudaMemcpyAsync(n_rays, d_nr, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRT_OK;
}

int trt_sample_batch(trt_ctx* c, trt_scene* s, const trt_sample_point* in, size_t n, const trt_params* p, trt_sampled_color* out) {
    if (!c || !s || !p || (n && (!in || !out))) return fail(TRT_ERR_INVALID, "trt_sample_batch: null argument");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, "trt_sample_batch: scene not committed (call trt_scene_commit)");
    static_assert(sizeof(trt_sample_point) == 32 && sizeof(trt_sampled_color) == 20, "wire structs must match the reference's repr(C) sizes");
    CU(cudaSetDevice(c->device));
    int rc; if ((rc = trt_scene_upload(c, s))) return rc;
    const trtk::SceneView& v = s->on_device[c->device].view;
    void *d_in = nullptr, *d_out = nullptr;
    if ((rc = ensure_scratch(c, 0, n * sizeof(trt_sample_point), &d_in))) return rc;
    if ((rc = ensure_scratch(c, 1, n * sizeof(trt_sampled_color), &d_out))) return rc;
    if (n) CU(cudaMemcpyAsync(d_in, in, n * sizeof(trt_sample_point), cudaMemcpyHostToDevice, c->stream));
    CU(trtk::launch_sample_batch(v, *p, (const trt_sample_point_pod*)d_in, n, (trt_sampled_color_pod*)d_out, c->stream));
    if (n) CU(cudaMemcpyAsync(out, d_out, n * sizeof(trt_sampled_color), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRT_OK;
}

// ---- streaming Sampler adapter: two slots of pinned host + device buffers on two streams, so that the H2D copy of batch
//      i+1 and the D2H copy of batch i-1 overlap the kernel of batch i (the reference's MetalSampler keeps every command
//      buffer in flight and waits at the end, metal/sampler.rs:107-130; here at most two batches are in flight).
struct SamplerSlot {
    cudaStream_t stream = nullptr; cudaEvent_t done = nullptr;
    trt_sample_point* h_in = nullptr; trt_sampled_color* h_out = nullptr; void* d_in = nullptr; void* d_out = nullptr;
    size_t n = 0; bool busy = false;
};
struct trt_sampler {
    trt_ctx* ctx; trt_scene* scene; trt_params params; size_t capacity; uint32_t next_sample;
    uint64_t scene_version;          // host scene version the device copy had when it was last (re-)uploaded for this adapter
    SamplerSlot slot[2]; int head = 0, tail = 0, in_flight = 0;   // FIFO over the two slots
};

int trt_sampler_create(trt_ctx* c, trt_scene* s, const trt_params* p, size_t batch_capacity, trt_sampler** out) {
    if (!c || !s || !p || !out || batch_capacity == 0) return fail(TRT_ERR_INVALID, "trt_sampler_create: bad argument");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, "trt_sampler_create: scene not committed (call trt_scene_commit)");
    CU(cudaSetDevice(c->device));
    int rc; if ((rc = trt_scene_upload(c, s))) return rc;
    trt_sampler* a = new (std::nothrow) trt_sampler();
    if (!a) return fail(TRT_ERR_OOM, "trt_sampler_create: host allocation failed");
    a->ctx = c; a->scene = s; a->params = *p; a->capacity = batch_capacity; a->next_sample = p->sample_offset;
    a->scene_version = s->host.version;
    for (SamplerSlot& k : a->slot) {
        cudaError_t e = cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&k.h_in, batch_capacity * sizeof(trt_sample_point));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&k.h_out, batch_capacity * sizeof(trt_sampled_color));
        if (e == cudaSuccess) e = cudaMalloc(&k.d_in, batch_capacity * sizeof(trt_sample_point));
        if (e == cudaSuccess) e = cudaMalloc(&k.d_out, batch_capacity * sizeof(trt_sampled_color));
        if (e != cudaSuccess) { trt_sampler_destroy(a); return fail_cuda("trt_sampler_create", e); }
    }
    *out = a;
    return TRT_OK;
}
int trt_sampler_destroy(trt_sampler* a) {
    if (!a) return TRT_OK;
    cudaSetDevice(a->ctx->device);
    for (SamplerSlot& k : a->slot) {
        if (k.stream) cudaStreamSynchronize(k.stream);
        if (k.h_in) cudaFreeHost(k.h_in);
        if (k.h_out) cudaFreeHost(k.h_out);
        if (k.d_in) cudaFree(k.d_in);
        if (k.d_out) cudaFree(k.d_out);
        if (k.done) cudaEventDestroy(k.done);
        if (k.stream) cudaStreamDestroy(k.stream);
    }
    delete a;
    return TRT_OK;
}
// Queues one batch (n <= capacity). Returns TRT_ERR_INVALID if both slots are in flight: collect first.
int trt_sampler_submit(trt_sampler* a, const trt_sample_point* in, size_t n) {
    if (!a || (n && !in)) return fail(TRT_ERR_INVALID, "trt_sampler_submit: null argument");
    if (n > a->capacity) return fail(TRT_ERR_INVALID, "trt_sampler_submit: batch larger than the adapter's capacity");
    if (a->in_flight == 2) return fail(TRT_ERR_INVALID, "trt_sampler_submit: two batches in flight, call trt_sampler_collect first");
    CU(cudaSetDevice(a->ctx->device));
    if (!a->scene->host.committed) return fail(TRT_ERR_INVALID, "trt_sampler_submit: the scene was modified and not re-committed");
    if (a->scene_version != a->scene->host.version || a->scene->on_device[a->ctx->device].version != a->scene->host.version) {
        // the scene was re-committed since the last batch: bring the device copy up to date (trt_scene_upload waits for every
        // stream of the device, so batches still in flight finish on the old copy first) and carry on with the new one
        int rc = trt_scene_upload(a->ctx, a->scene); if (rc) return rc;
        a->scene_version = a->scene->host.version;
    }
    SamplerSlot& k = a->slot[a->head];
    const trtk::SceneView& v = a->scene->on_device[a->ctx->device].view;
    if (n) std::memcpy(k.h_in, in, n * sizeof(trt_sample_point));
    trt_params p = a->params; p.sample_offset = a->next_sample;        // item i of the whole stream uses sample_offset + i
    if (n) CU(cudaMemcpyAsync(k.d_in, k.h_in, n * sizeof(trt_sample_point), cudaMemcpyHostToDevice, k.stream));
    CU(trtk::launch_sample_batch(v, p, (const trt_sample_point_pod*)k.d_in, n, (trt_sampled_color_pod*)k.d_out, k.stream));
    if (n) CU(cudaMemcpyAsync(k.h_out, k.d_out, n * sizeof(trt_sampled_color), cudaMemcpyDeviceToHost, k.stream));
    CU(cudaEventRecord(k.done, k.stream));
    k.n = n; k.busy = true;
    a->next_sample += (uint32_t)n; a->head ^= 1; a->in_flight++;
    return TRT_OK;
}
// Waits for the oldest batch in flight and copies its colours out (same order as submitted). *n_out = 0 when nothing is in flight.
int trt_sampler_collect(trt_sampler* a, trt_sampled_color* out, size_t* n_out) {
    if (!a || !n_out) return fail(TRT_ERR_INVALID, "trt_sampler_collect: null argument");
    *n_out = 0;
    if (a->in_flight == 0) return TRT_OK;
    CU(cudaSetDevice(a->ctx->device));
    SamplerSlot& k = a->slot[a->tail];
    CU(cudaEventSynchronize(k.done));
    if (k.n && !out) return fail(TRT_ERR_INVALID, "trt_sampler_collect: out is null");
    if (k.n) std::memcpy(out, k.h_out, k.n * sizeof(trt_sampled_color));
    *n_out = k.n; k.busy = false;
    a->tail ^= 1; a->in_flight--;
    return TRT_OK;
}

int trt_trace_rays(trt_ctx* c, trt_scene* s, size_t n, const float* origins, const float* dirs, float t_min, uint32_t* prim_id, float* t) {
    if (!c || !s || (n && (!origins || !dirs || !prim_id || !t))) return fail(TRT_ERR_INVALID, "trt_trace_rays: null argument");
    if (!s->host.committed) return fail(TRT_ERR_INVALID, "trt_trace_rays: scene not committed (call trt_scene_commit)");
    CU(cudaSetDevice(c->device));
    int rc; if ((rc = trt_scene_upload(c, s))) return rc;
    const trtk::SceneView& v = s->on_device[c->device].view;
    void *d_o = nullptr, *d_d = nullptr, *d_p = nullptr, *d_t = nullptr;
    if ((rc = ensure_scratch(c, 0, n * 12, &d_o))) return rc;
    if ((rc = ensure_scratch(c, 1, n * 12, &d_d))) return rc;
    if ((rc = ensure_scratch(c, 2, n * 4, &d_p))) return rc;
    if ((rc = ensure_scratch(c, 3, n * 4, &d_t))) return rc;
    if (n) { CU(cudaMemcpyAsync(d_o, origins, n * 12, cudaMemcpyHostToDevice, c->stream)); CU(cudaMemcpyAsync(d_d, dirs, n * 12, cudaMemcpyHostToDevice, c->stream)); }
    CU(trtk::launch_trace(v, (const float*)d_o, (const float*)d_d, n, t_min, (uint32_t*)d_p, (float*)d_t, c->stream));
    if (n) { CU(cudaMemcpyAsync(prim_id, d_p, n * 4, cudaMemcpyDeviceToHost, c->stream)); CU(cudaMemcpyAsync(t, d_t, n * 4, cudaMemcpyDeviceToHost, c->stream)); }
    CU(cudaStreamSynchronize(c->stream));
    return TRT_OK;
}

int trt_debug_math(trt_ctx* c, size_t n, const float* u, float* s, float* co, float* cb) {
    if (!c || (n && (!u || !s || !co || !cb))) return fail(TRT_ERR_INVALID, "trt_debug_math: null argument");
    CU(cudaSetDevice(c->device));
    int rc; void* d[4];
    for (int i = 0; i < 4; ++i) if ((rc = ensure_scratch(c, i, n * 4, &d[i]))) return rc;
    if (n) CU(cudaMemcpyAsync(d[0], u, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(trtk::launch_debug_math(n, (const float*)d[0], (float*)d[1], (float*)d[2], (float*)d[3], c->stream));
    if (n) {
        CU(cudaMemcpyAsync(s, d[1], n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(co, d[2], n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(cb, d[3], n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return TRT_OK;
}
int trt_debug_philox(trt_ctx* c, size_t n, const uint32_t* ctr, const uint32_t key[2], uint32_t* out) {
    if (!c || !key || (n && (!ctr || !out))) return fail(TRT_ERR_INVALID, "trt_debug_philox: null argument");
    CU(cudaSetDevice(c->device));
    int rc; void *d_c = nullptr, *d_o = nullptr;
    if ((rc = ensure_scratch(c, 0, n * 16, &d_c))) return rc;
    if ((rc = ensure_scratch(c, 1, n * 16, &d_o))) return rc;
    if (n) CU(cudaMemcpyAsync(d_c, ctr, n * 16, cudaMemcpyHostToDevice, c->stream));
    CU(trtk::launch_debug_philox(n, (const uint32_t*)d_c, key[0], key[1], (uint32_t*)d_o, c->stream));
    if (n) CU(cudaMemcpyAsync(out, d_o, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TRT_OK;
}

} // extern "C"
