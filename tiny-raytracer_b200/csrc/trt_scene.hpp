// trt_scene.hpp — host side of the scene: the World registry (hittable/world.rs), the two BVH builders and the
// flattener that produces the HBM layout of trt_layout.h. Plain C++17, no CUDA.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "trt_layout.h"

namespace trt {

struct HostMaterial { uint32_t kind; float rgb[3]; float param; };

struct HostPrim {
    uint32_t type;        // 0 sphere, 1 quad
    uint32_t mat;
    float a[3], b[3], c[3]; float r;   // sphere: a = center, r = radius; quad: a = corner, b = u, c = v
    float n[3], w[3], d;               // quad derived (hittable/quad.rs:25-27)
    float bmin[3], bmax[3];            // padded AABB (hittable/aabb.rs:13-20)
    uint32_t type_index;               // index inside its type's array
};

struct FlatScene {
    std::vector<trt_f4> nodes;         // 4 per interior node
    std::vector<trt_f4> nodes4;        // TRT_BVH_SAH: the same tree collapsed to 4-wide nodes, 8 per node (trt_scene.cpp collapse4); may be empty
    uint32_t n_nodes4 = 0, max_depth4 = 0;
    std::vector<trt_f4> nodes_ref4;    // TRT_BVH_SAH: wide nodes of the reference-topology tree, walked only to replay order-dependent rays
    uint32_t n_nodes_ref4 = 0;
    bool wide_on_device = false;       // ... and so are its own wide nodes (plain two-level collapse of the uploaded binary nodes)
    bool ref_on_device = false;        // TRT_BVH_SAH with many primitives: ranks and nodes_ref4 are produced during trt_scene_upload (lbvh_boxes is their input)
    std::vector<trt_f4> spheres;
    std::vector<trt_f4> quads;         // 4 per quad
    std::vector<trt_u4> sphere_meta, quad_meta;
    std::vector<trt_f4> quad_nn;       // per quad: n / |n| (what HitRecord::new computes at every hit, hittable/mod.rs:36-40), same IEEE ops
    std::vector<trt_f4> materials;     // 2 per material
    std::vector<trt_f4> flat;          // TRT_BVH_FLAT: TRT_FLAT_F4 per primitive, reference leaf order (trt_layout.h)
    uint32_t flat_qmask[2] = {0u, 0u}; // TRT_BVH_FLAT: bit r = the record of leaf rank r is a quad (uniform type lookup without a load)
    std::vector<trt_f4> flat_cbox;     // TRT_BVH_FLAT: conservative centre/half-width boxes for phase A (trt_device.cuh), 3 per pair, NaN padded
    float flat_bound = -1.f;           // largest |coordinate| of those boxes; < 0 = phase A disabled (coordinates beyond 2^30 or not finite)
    std::vector<trt_f4> lbvh_boxes;    // TRT_BVH_LBVH_DEVICE: 2 per primitive, (bmin, leaf-ref bits), (bmax, 0): input of the device builder
    uint32_t root = TRT_REF_INVALID;   // interior node 0, or INVALID for an empty scene
    uint32_t n_nodes = 0, max_depth = 0;
    float sah_cost = 0.f;
    int bvh_kind = 0;
    // back to the empty state WITHOUT giving the arrays' memory back: a re-commit of a big scene then writes into pages it
    // already owns (10^6 primitives: 64 MB of fresh page faults per commit otherwise)
    void reset(int kind) {
        nodes.clear(); nodes4.clear(); nodes_ref4.clear(); spheres.clear(); quads.clear(); sphere_meta.clear(); quad_meta.clear();
        quad_nn.clear(); materials.clear(); flat.clear(); flat_cbox.clear(); lbvh_boxes.clear();
        n_nodes4 = max_depth4 = n_nodes_ref4 = 0; wide_on_device = ref_on_device = false; flat_qmask[0] = flat_qmask[1] = 0u;
        flat_bound = -1.f; root = TRT_REF_INVALID; n_nodes = max_depth = 0; sah_cost = 0.f; bvh_kind = kind;
    }
    size_t bytes() const {
        return (nodes.size() + spheres.size() + quads.size() + materials.size()) * sizeof(trt_f4) +
               (sphere_meta.size() + quad_meta.size()) * sizeof(trt_u4);
    }
};

struct HostScene {
    std::vector<HostMaterial> materials;
    std::vector<HostPrim> prims;        // insertion order == prim_id
    std::vector<uint32_t> leaf_rank;    // reference-topology leaf order per prim_id (after commit)
    FlatScene flat;
    bool committed = false;
    float host_commit_ms = 0.f;         // wall time of the last commit (host tree builds + flattening)
    float device_build_ms = 0.f;        // TRT_BVH_LBVH_DEVICE: time of the last build on the GPU (CUDA events)
    uint64_t version = 0;               // bumped by every commit; the device copy remembers which one it holds

    uint32_t add_material(uint32_t kind, const float rgb[3], float param);
    // returns false + message on bad input
    bool add_sphere(const float c[3], float r, uint32_t mat, uint32_t* id, std::string* err);
    bool add_quad(const float corner[3], const float u[3], const float v[3], uint32_t mat, uint32_t* id, std::string* err);
    bool commit(int bvh_kind, std::string* err);
};

} // namespace trt
