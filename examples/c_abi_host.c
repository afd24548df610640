/* examples/c_abi_host.c — include/trt.h from plain C99: the host-only half of the boundary (scene, commit, camera) needs no GPU.
 * Build: gcc -std=c99 -pedantic -Iinclude examples/c_abi_host.c -o c_abi_host -Ltiny-raytracer_b200 -ltrt_b200 */
#include <stdio.h>
#include "trt.h"

int main(void) {
    trt_scene* s = 0; trt_camera cam; trt_params p; unsigned id = 0, q = 0, n = 0, nodes = 0, depth = 0;
    float white[3] = {0.7f, 0.7f, 0.7f}, c[3] = {0, 0, -1}, c2[3] = {0, -100.5f, -1}, up[3] = {0, 1, 0}, pos[3] = {0, 0, 1}, look[3] = {0, 0, 0};
    float cost = 0.f;
    if (trt_scene_create(&s)) return 2;
    if (trt_add_lambertian(s, white, &id) || trt_add_sphere(s, c, 0.5f, id, &q) || trt_add_sphere(s, c2, 100.0f, id, &q)) return 3;
    if (trt_add_sphere(s, c, 0.5f, 77u, &q) == TRT_OK) return 4;          /* unknown material: an error value, not a crash */
    if (trt_scene_commit(s, TRT_BVH_SAH)) return 5;
    if (trt_scene_num_prims(s, &n) || n != 2 || trt_scene_bvh_info(s, &nodes, &depth, &cost) || nodes != 1) return 6;
    if (trt_camera_make(1.0f, 0.0f, pos, look, up, 60.0f, 16, 8, &cam) || cam.width != 16) return 7;
    p.width = 16; p.height = 8; p.spp = 1; p.sample_offset = 0; p.max_bounces = 4; p.seed = 1; p.t_min = 0.001f; p.spp_total = 0;
    p.variant = TRT_VARIANT_MEGAKERNEL; p.flags = TRT_FLAG_NO_TAIL_SPLIT | TRT_FLAG_REGEN_MIN(8);
    p.background[0] = p.background[1] = p.background[2] = 1.0f;
    printf("abi %u, %u prims, %u node(s), sample point %u bytes, flags %u, last error: %s\n", trt_abi_version(), n, nodes,
           (unsigned)sizeof(trt_sample_point), p.flags, trt_last_error());
    return trt_scene_destroy(s);
}
