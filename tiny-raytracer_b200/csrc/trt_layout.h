// trt_layout.h — HBM data layout shared by the host flattener (trt_scene.cpp) and the kernels (trt_kernels.cu).
//
// Everything is 16-byte-vector addressable; a node is one 64-byte record (four float4) so that one visited interior
// node costs four 128-bit loads from one 64-byte-aligned segment (half a 128-byte line):
//
//   node n (float4 nodes[4n .. 4n+3]):
//     [0] = ( lmin.x, rmin.x, lmin.y, rmin.y )      the two children's padded AABBs (hittable/aabb.rs:5-10 is 24 B {min,max}),
//     [1] = ( lmin.z, rmin.z, lmax.x, rmax.x )      interleaved as (left, right) pairs per coordinate: each 64-bit half of a
//     [2] = ( lmax.y, rmax.y, lmax.z, rmax.z )      load is one operand of a packed FADD2/FMUL2 (both slab tests at once)
//     [3] = ( left_ref, right_ref, 0, 0 ) as bits
//   wide node n (float4 nodes4[8n .. 8n+7], TRT_BVH_SAH and the device-built tree): [0..2] children 0,1 and [3..5] children 2,3
//     in the same pair layout, [6] = the four child refs, [7] unused; absent child = TRT_REF_INVALID with an empty box.
//   child ref: bit31 = leaf, bit30 = primitive type (0 sphere, 1 quad), bits0..29 = index into that type's array;
//              otherwise the index of an interior node; TRT_REF_INVALID = absent child (single-primitive scene).
//   With one primitive per leaf the child box stored in the parent IS the primitive's own padded box, i.e. the box the
//   reference tests in the leaf's Node::hit (hittable/bvh.rs:89) — no extra per-primitive box fetch.
//
//   sphere i : float4 spheres[i]      = ( center.xyz, radius )                                    16 B (hittable/sphere.rs:8-13)
//   quad i   : float4 quads[4i..4i+3] = ( n.xyz, d ), ( corner.xyz, w.x ), ( u.xyz, w.y ), ( v.xyz, w.z )   64 B (hittable/quad.rs:8-17)
//              the plane test needs only the first 16 B; the rest is fetched when t is in range.
//   meta     : uint4 sphere_meta[i] / quad_meta[i] = ( material_id, reference_leaf_rank, prim_id, 0 ), read for the final hit
//              (and on exact-t ties in SAH mode).
//   material m: float4 materials[2m] = ( rgb, kind as bits ), materials[2m+1] = ( fuzz | refraction_index, 0, 0, 0 )   32 B
#pragma once
#include <stdint.h>

#define TRT_REF_INVALID 0x7FFFFFFFu
#define TRT_REF_LEAF 0x80000000u
#define TRT_REF_QUAD 0x40000000u
#define TRT_REF_INDEX_MASK 0x3FFFFFFFu

#define TRT_MAT_LAMBERTIAN 0u
#define TRT_MAT_METAL 1u
#define TRT_MAT_DIELECTRIC 2u
#define TRT_MAT_LIGHT 3u

#define TRT_MAX_STACK 64   /* traversal stack entries; trt_scene_commit rejects deeper trees */
#define TRT_AUTO_DEVICE_BUILD_MIN 65536u   /* TRT_BVH_AUTO: from this many primitives on the trees are built on the GPU
                                               (10^6 spheres: 5 ms instead of 1.1 s on the host, traversal -3 %, same results) */
#define TRT_FLAT_MAX 64    /* TRT_BVH_FLAT: largest primitive count handled as a flat list */
#ifndef TRT_FLAT_F4
#define TRT_FLAT_F4 7
#endif
                           /* float4 per flat-list record (112 B): the four geometry float4 above, then the primitive's own padded
                              box as (min, max) pairs (mn.x, mx.x, mn.y, mx.y), (mn.z, mx.z, child-ref bits, 0) — each 64-bit half
                              is one operand of a packed FADD2 / FMUL2 in the exact box test —, then one spare. An ODD number of
                              16-byte slots, so that lanes reading the same field of different records hit different
                              shared-memory bank groups (96-byte records put every other record on the same banks: 37 % of
                              the flat kernel's shared-memory wavefronts were bank conflicts, profiles/r1y) */

struct trt_f4 { float x, y, z, w; };
struct trt_u4 { uint32_t x, y, z, w; };
