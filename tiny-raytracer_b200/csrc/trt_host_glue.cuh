// trt_host_glue.cuh — host-side conversion of the C-ABI structs into the kernels' argument structs (shared by the
// megakernel and wavefront translation units).
#pragma once
#include <cmath>
#include "trt_device.cuh"
#include "trt_kernels.h"

namespace trtk {
using namespace trtd;

inline DevScene dev_scene(const SceneView& v) {
    DevScene s;
    s.nodes = (const float4*)v.nodes; s.spheres = (const float4*)v.spheres; s.quads = (const float4*)v.quads;
    s.sphere_meta = (const uint4*)v.sphere_meta; s.quad_meta = (const uint4*)v.quad_meta; s.quad_nn = (const float4*)v.quad_nn; s.materials = (const float4*)v.materials;
    s.flat = (const FlatPrim*)v.flat; s.n_flat = v.n_flat; s.flat_cbox = (const float4*)v.flat_cbox; s.flat_bound = v.flat_bound; s.flat_qmask[0] = v.flat_qmask[0]; s.flat_qmask[1] = v.flat_qmask[1];
    s.nodes4 = (const float4*)v.nodes4; s.n_nodes4 = v.n_nodes4;
    s.nodes_ref4 = (const float4*)v.nodes_ref4; s.n_nodes_ref4 = v.n_nodes_ref4; s.root_ref4 = v.root_ref4; s.g_spheres = s.spheres; s.g_quads = s.quads;
    s.nodes8 = (const float4*)v.nodes8; s.n_nodes8 = v.n_nodes8; s.quad_box = (const float4*)v.quad_box; s.bound8 = v.bound8;
    s.root = v.root; s.n_nodes = v.n_nodes; s.n_spheres = v.n_spheres; s.n_quads = v.n_quads; s.n_materials = v.n_materials;
    s.ordered = v.kind == 1 ? 1u : 0u;
    for (int i = 0; i < 3; ++i) { s.root_min[i] = v.root_min[i]; s.root_max[i] = v.root_max[i]; }
    return s;
}
inline DevCamera dev_camera(const trt_camera& c) {
    DevCamera d;
    d.position = mk(c.position[0], c.position[1], c.position[2]);
    d.upper_left = mk(c.viewport_upper_left[0], c.viewport_upper_left[1], c.viewport_upper_left[2]);
    d.horizontal = mk(c.horizontal[0], c.horizontal[1], c.horizontal[2]);
    d.vertical = mk(c.vertical[0], c.vertical[1], c.vertical[2]);
    d.du = mk(c.defocus_disk_u[0], c.defocus_disk_u[1], c.defocus_disk_u[2]);
    d.dv = mk(c.defocus_disk_v[0], c.defocus_disk_v[1], c.defocus_disk_v[2]);
    d.width = c.width; d.height = c.height;
    bool pin = true;
    for (int i = 0; i < 3; ++i) {
        if (c.defocus_disk_u[i] != 0.0f || c.defocus_disk_v[i] != 0.0f) pin = false;
        if (c.position[i] == 0.0f && std::signbit(c.position[i])) pin = false;     // -0 + (+0) = +0 but -0 + (-0) = -0
    }
    d.pinhole = pin ? 1u : 0u;
    return d;
}
inline DevParams dev_params(const trt_params& p) {
    DevParams d;
    d.width = p.width; d.height = p.height; d.spp = p.spp; d.sample_offset = p.sample_offset; d.max_bounces = p.max_bounces;
    d.bg[0] = p.background[0]; d.bg[1] = p.background[1]; d.bg[2] = p.background[2];
    d.seed_lo = (uint32_t)p.seed; d.seed_hi = (uint32_t)(p.seed >> 32); d.t_min = p.t_min;
    return d;
}


} // namespace trtk
