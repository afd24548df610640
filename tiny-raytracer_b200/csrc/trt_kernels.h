// trt_kernels.h — host-callable launchers of trt_kernels.cu (all pointers are DEVICE pointers).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/trt.h"

struct trt_sample_point_pod { uint32_t x, y; float o[3], d[3]; };   // == trt_sample_point
struct trt_sampled_color_pod { uint32_t x, y; float c[3]; };        // == trt_sampled_color

namespace trtk {

struct SceneView {   // device copy of trt::FlatScene
    const void* nodes; const void* spheres; const void* quads; const void* sphere_meta; const void* quad_meta; const void* materials; const void* quad_nn;
    const void* flat;          // device copy of the flat list (TRT_FLAT_F4 float4 per primitive, reference leaf order) or null
    uint32_t flat_qmask[2];    // bit r = record r of the flat list is a quad
    const void* flat_cbox;     // device copy of the flat list's conservative phase-A boxes (TRT_FLAT_CBOX_F4 float4) or null
    const void* flat_cbox_host;// the same boxes on the host: copied into the kernel parameters of k_render_mega_flat
    float flat_bound;          // largest |coordinate| of any flat-list box, or < 0: phase A disabled (literal scan only)
    uint32_t n_flat;
    const void* nodes4; uint32_t n_nodes4;   // tree kinds: the tree collapsed to 4-wide nodes (8 float4 each) or null
    const void* nodes_ref4; uint32_t n_nodes_ref4, root_ref4;   // TRT_BVH_SAH: wide nodes of the reference-topology tree (replay of order-dependent rays) or null
    const void* nodes8; uint32_t n_nodes8;   // ordered kinds in global memory: compressed 8-wide nodes (6 float4 each, root 0) or null
    const void* quad_box;                    // with nodes8: every quad's own padded box, (min.xyz, 0), (max.xyz, 0)
    float bound8;                            // with nodes8: largest |coordinate| of the scene's boxes
    uint32_t root, n_nodes, n_spheres, n_quads, n_materials;
    uint32_t kind;             // trt_bvh_kind after AUTO resolution: 0 reference, 1 SAH, 2 flat
    float root_min[3], root_max[3];
};

// kind as walked: 3 / 4 = the 4-wide nodes of a TRT_BVH_SAH / TRT_BVH_REFERENCE scene (when it has them and binary_only is false)
uint32_t walk_kind(const SceneView& v, bool binary_only);
size_t scene_smem_bytes(const SceneView& v, uint32_t kind);

// d_primary: null, or primary_record_bytes(p) * p.spp bytes of scratch: the coherent first-bounce pass (k_primary) then runs
// before the path loop, which starts every path from its record. Same results either way.
size_t primary_record_bytes(const trt_params& p);
bool primary_pass_pays(const SceneView& v, const trt_params& p);     // the configurations where the pass was measured to be faster
cudaError_t launch_render_mega(const SceneView& v, const trt_camera& cam, const trt_params& p, double* d_sum64,
                               unsigned long long* d_counters, unsigned int* d_next_item, int sm_count, cudaStream_t st,
                               void* d_primary);
size_t lbvh_tmp_bytes(uint32_t n);
// Builds the 64-byte node array (n - 1 nodes) and, if d_nodes4 is not null, its 4-wide form (128 bytes per node) on the device from per-primitive padded boxes (2 float4 each, leaf-ref bits in bmin.w).
cudaError_t build_lbvh(const void* d_boxes, uint32_t n, void* d_nodes, void* d_nodes4, void* d_tmp, uint32_t* h_depth, float* h_ms, cudaStream_t st);
cudaError_t collapse_lbvh4(const void* d_nodes, uint32_t n, void* d_nodes4, bool greedy, cudaStream_t st);
// Compressed 8-wide nodes (trt_lbvh.cu k_cw8_level) of a binary node array with root `root`: 96 bytes per wide node into d_nodes8
// (room for n - 1 of them), their number into *h_n_nodes8. Synchronises the stream.
size_t cw8_tmp_bytes(uint32_t n);
cudaError_t build_cw8(const void* d_nodes, uint32_t n, uint32_t root, void* d_nodes8, void* d_tmp, uint32_t* h_n_nodes8, uint32_t* h_levels, cudaStream_t st);
size_t ref_order_tmp_bytes(uint32_t n);
// Node::new (hittable/bvh.rs:42-84) on the device: the reference leaf rank of every primitive into meta.y and the wide nodes
// of the reference-topology tree (8 float4 per node, n - 1 nodes, root index in *h_root) into d_nodes_ref4.
cudaError_t build_reference_order(const void* d_boxes, uint32_t n, void* d_sphere_meta, void* d_quad_meta, void* d_nodes_ref4,
                                  void* d_tmp, uint32_t* h_root, float* h_ms, cudaStream_t st);
size_t wavefront_scratch_bytes(uint32_t n_slots);
// Runs the whole wavefront render (host loop inside; synchronises the stream). h_rays_paths[0] = rays, [1] = paths.
cudaError_t launch_render_wavefront(const SceneView& v, const trt_camera& cam, const trt_params& p, double* d_sum64,
                                    void* d_scratch, uint32_t n_slots, void* d_ctl, int sm_count, cudaStream_t st,
                                    unsigned long long* h_rays_paths, uint32_t* launches_out);
cudaError_t launch_aov(const SceneView& v, const trt_camera& cam, const trt_params& p, uint32_t* d_prim, float* d_t,
                       float* d_rad, uint32_t* d_nrays, cudaStream_t st);
cudaError_t launch_sample_batch(const SceneView& v, const trt_params& p, const trt_sample_point_pod* d_in, size_t n,
                                trt_sampled_color_pod* d_out, cudaStream_t st);
cudaError_t launch_trace(const SceneView& v, const float* d_o, const float* d_d, size_t n, float t_min, uint32_t* d_prim,
                         float* d_t, cudaStream_t st);
cudaError_t launch_finalize(const double* d_sum64, size_t n_values, uint32_t spp_total, float* d_linear, uint8_t* d_rgb8, cudaStream_t st);
// Multi-GPU exchange + epilogue (trt_multi.cu): slice [begin, end) of the image = sum over d_src[0..n_src) in rank order, normalised
// and tone-mapped. d_src are device pointers valid on the launching device (own memory or peer-mapped memory of the other GPUs).
#define TRT_MULTI_MAX_DEVICES 16
cudaError_t launch_reduce_scatter_finalize(const double* const* d_src, int n_src, size_t begin, size_t end, uint32_t spp_total,
                                           float* d_lin, uint8_t* d_rgb8, int sm_count, cudaStream_t st);
cudaError_t launch_debug_math(size_t n, const float* u, float* s, float* c, float* cb, cudaStream_t st);
cudaError_t launch_debug_philox(size_t n, const uint32_t* ctr, uint32_t k0, uint32_t k1, uint32_t* out, cudaStream_t st);

} // namespace trtk
