// trt_device.cuh — device-side arithmetic of the path-tracing hot path for sm_100a.
//
// Numeric contract: every f32 operation below is a separately rounded IEEE-754 operation in the
// order the reference performs it (Rust never contracts a*b+c). The translation unit is compiled with
// -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, so `*`,`+`,`/`,sqrtf are mul.rn/add.rn/div.rn/sqrt.rn.
// Citations are file:line relative to /root/reference/raytracer/src.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "trt_layout.h"

// -DTRT_DEBUG_BOUNDS: device-side asserts on every traversal-stack push and node / primitive index (a debug build for the
// GPU box; compute-sanitizer is not available on this pool). Off in the product build.
#ifdef TRT_DEBUG_BOUNDS
#include <assert.h>
#define TRT_ASSERT(c) assert(c)
#else
#define TRT_ASSERT(c) ((void)0)
#endif

namespace trtd {

#define TRT_DEV __device__ __forceinline__
#define TRT_INF __int_as_float(0x7f800000)

struct V3 { float x, y, z; };
__host__ __device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
TRT_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }          // math/vec3.rs:70-76
TRT_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }          // math/vec3.rs:86-92
TRT_DEV V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
TRT_DEV V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }             // math/vec3.rs:102-108
TRT_DEV V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }             // math/vec3.rs:110-116
TRT_DEV V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }          // math/vec3.rs:126-132
TRT_DEV V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }             // math/vec3.rs:142-148
TRT_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }               // math/vec3.rs:49-51
TRT_DEV V3 cross(V3 a, V3 b) {                                                            // math/vec3.rs:53-59
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
TRT_DEV V3 normalized(V3 a) { return a / sqrtf(dot(a, a)); }                              // math/vec3.rs:45-47
TRT_DEV bool near_zero(V3 a) { const float e = 1e-7f; return fabsf(a.x) < e && fabsf(a.y) < e && fabsf(a.z) < e; } // math/vec3.rs:32-35
TRT_DEV V3 reflect(V3 v, V3 n) { return v - (2.0f * dot(v, n)) * n; }                     // math/vec3extend.rs:75-77
TRT_DEV V3 refract(V3 v, V3 n, float eta) {                                               // math/vec3extend.rs:79-84
    float c = fminf(-dot(n, v), 1.0f);
    V3 perp = eta * (v + n * c);
    V3 par = (-sqrtf(fabsf(1.0f - dot(perp, perp)))) * n;
    return par + perp;
}
TRT_DEV V3 ray_at(V3 o, V3 d, float t) { return o + t * d; }                              // ray.rs:24-26

// ------------------------------------------------------------------ Philox4x32-10, key (seed), counter (pixel, sample, bounce, block)
struct U4 { uint32_t x, y, z, w; };
TRT_DEV U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3; return o;
}
// rand 0.8.5 UniformFloat<f32>: 23 mantissa bits -> [0,1)
TRT_DEV float u01(uint32_t b) { return (float)(b >> 9) * 1.1920928955078125e-07f; }
TRT_DEV float u11(uint32_t b) { return u01(b) * 2.0f + -1.0f; }   // random_range(-1.0..1.0): f*(hi-lo)+lo

// ------------------------------------------------------------------ deterministic sin/cos(2*pi*u), cbrt
TRT_DEV void det_sincos2pi(float u, float* s_out, float* c_out) {
    float q = u * 8.0f;
    int k = (int)q;
    float f = q - (float)k;
    float r = (k & 1) ? (1.0f - f) : f;
    float x = r * 0.785398163397448309616f;
    float z = x * x;
    float sp = -1.9515295891e-4f * z;
    sp = sp + 8.3321608736e-3f;
    sp = sp * z;
    sp = sp - 1.6666654611e-1f;
    sp = sp * z;
    sp = sp * x;
    float s = sp + x;
    float cp = 2.443315711809948e-5f * z;
    cp = cp - 1.388731625493765e-3f;
    cp = cp * z;
    cp = cp + 4.166664568298827e-2f;
    cp = cp * z;
    cp = cp * z;
    float hz = 0.5f * z;
    cp = cp - hz;
    float c = cp + 1.0f;
    float ss = s, cc = c;
    if ((k + 1) & 2) { ss = c; cc = s; }
    if (k >= 4) ss = -ss;
    if (k >= 2 && k <= 5) cc = -cc;
    *s_out = ss; *c_out = cc;
}
TRT_DEV float det_cbrt(float x) {
    if (!(x > 0.0f)) return 0.0f;
    uint32_t b = __float_as_uint(x);
    b = b / 3u + 709921077u;
    float y = __uint_as_float(b);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float y2 = y * y;
        float y3 = y2 * y;
        float num = y3 - x;
        float den = 3.0f * y2;
        y = y - num / den;
    }
    return y;
}
// math/vec3extend.rs:15-30 with the contract's substitutions (cos(phi) = 1-2u2, sin(phi) = sqrt(1-cos^2))
TRT_DEV V3 random_in_unit_sphere(float u1, float u2, float u3, bool use_r) {
    float sin_t, cos_t;
    det_sincos2pi(u1, &sin_t, &cos_t);
    float cos_phi = 1.0f - 2.0f * u2;
    float s2 = 1.0f - cos_phi * cos_phi;
    float sin_phi = sqrtf(fmaxf(s2, 0.0f));
    float r = use_r ? det_cbrt(u3) : 1.0f;
    return mk(r * sin_phi * cos_t, r * sin_phi * sin_t, r * cos_phi);
}

// ------------------------------------------------------------------ scene access
struct FlatPrim;
struct DevScene {
    const float4* nodes; const float4* spheres; const float4* quads;
    const uint4* sphere_meta; const uint4* quad_meta; const float4* materials;
    const float4* quad_nn;                      // unit normal of every quad, precomputed on the host with HitRecord::new's operations
    const FlatPrim* flat; uint32_t n_flat;      // TRT_BVH_FLAT: primitives in reference leaf order (else null / 0)
    uint32_t flat_qmask[2];                     // TRT_BVH_FLAT: bit r = record r is a quad
    const float4* nodes4; uint32_t n_nodes4;    // tree kinds: 4-wide nodes (8 float4 each, trt_scene.cpp collapse4) or null
    const float4* nodes_ref4; uint32_t n_nodes_ref4, root_ref4;   // ordered kinds: the REFERENCE-topology tree's wide nodes, for the rare replay (ordered_replay)
    const float4* g_spheres; const float4* g_quads;    // global-memory copies of spheres / quads (the staged scene overrides those two)
    const float4* flat_cbox; float flat_bound;  // TRT_BVH_FLAT: conservative centre/half-width boxes (phase A), largest |coordinate|
    const float4* nodes8; uint32_t n_nodes8;    // ordered kinds in global memory: compressed 8-wide nodes (6 float4 each, root 0; trt_lbvh.cu k_cw8_level) or null
    const float4* quad_box; float bound8;       // with nodes8: every quad's own padded box (2 float4), largest |coordinate| of the scene's boxes
    uint32_t root, n_nodes, n_spheres, n_quads, n_materials, ordered;
    float root_min[3], root_max[3];   // the root's own padded box (only consulted when counting events)
};
struct DevCamera { V3 position, upper_left, horizontal, vertical, du, dv; uint32_t width, height;
                   uint32_t pinhole; };   // du == dv == 0 and no -0 in position: the lens sample cannot change the origin
struct DevParams {
    uint32_t width, height, spp, sample_offset, max_bounces;
    float bg[3]; uint32_t seed_lo, seed_hi; float t_min;
};

// The per-hit records (meta, materials, quad normals) are read once per ray segment, through GENERIC loads: they sit in
// shared memory when they are small and stay in global memory (L1/L2) when staging them would cost occupancy.
TRT_DEV float4 ld4a(const float4* p) { return *p; }
TRT_DEV uint4 ld4ua(const uint4* p) { return *p; }
#ifndef TRT_LD4_QUAL
#define TRT_LD4_QUAL ""              /* A/B: cache qualifier (e.g. ".L1::no_allocate", see TRT_NODE_LD_QUAL) of the 128-bit loads of primitives / binary nodes in global memory */
#endif
// SMEM = scene staged in shared memory (small scenes): ld.shared.v4; otherwise read-only global path.
template <bool SMEM> TRT_DEV float4 ld4(const float4* p) {
    if (SMEM) {
        float4 v; uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        return v;
    } else if (sizeof(TRT_LD4_QUAL) > 1) {
        float4 v;
        asm("ld.global.nc" TRT_LD4_QUAL ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        return v;
    } else return __ldg(p);
}
template <bool SMEM> TRT_DEV uint4 ld4u(const uint4* p) {
    if (SMEM) {
        uint4 v; uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        return v;
    } else return __ldg(p);
}

// Two consecutive float4 of a 32-byte-aligned record in ONE 256-bit load (LDG.E.256, new with sm_100). A lane whose record lies
// in its own cache line costs the L1 data pipe one wavefront per SECTOR it touches per instruction (10^6 spheres:
// l1tex__data_pipe_lsu_wavefronts 92 % of peak, one per 16-byte lane access, profiles/r2d_tree_c4_64spp): fetching a wide node
// as 3 x 32 B + 16 B instead of 7 x 16 B takes the same bytes through 4 instead of 7 of them.
// Cache qualifier of the wide-node loads of trees in global memory. ".L1::no_allocate": the records still come through L1 but do
// not displace what is there — on 10^6 spheres L1 then holds the traversal stacks and spills (local memory, whose misses sit
// inside the pop chain) instead of node lines that are hardly ever reused by the same SM: 2 317 -> 2 511 Mrays/s (device-built
// tree), 2 451 -> 2 689 (SAH). Measured and rejected (profiles/r2h_*): ".L1::evict_last" +-0; letting the first 2 / 4 node
// fetches of every ray (the top of the tree) allocate -3 %; the same qualifier on the primitive loads -0.5 %; 5 CTAs per SM
// -6 %; an L2 access-policy window (persisting set-aside of 32..79 MB) on the node array -2.5 %.
#ifndef TRT_NODE_LD_QUAL
#define TRT_NODE_LD_QUAL ".L1::no_allocate"
#endif
struct F4x2 { float4 a, b; };
TRT_DEV F4x2 ldg256(const float4* p) {
    F4x2 v;
    asm("ld.global.nc" TRT_NODE_LD_QUAL ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w) : "l"(p));
    return v;
}
TRT_DEV float4 ldg128n(const float4* p) {      // the node's seventh float4, with the same qualifier
    float4 v;
    asm("ld.global.nc" TRT_NODE_LD_QUAL ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

struct EventCounters { uint32_t node_visits, sphere_tests, quad_tests; };

// ------------------------------------------------------------------ packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2)
// Each half is an ordinary IEEE round-to-nearest operation (denormals kept), so a packed add or mul gives exactly the bits
// of the two scalar instructions it replaces; only the issue slot is shared.
typedef unsigned long long f32x2;
TRT_DEV f32x2 pk2(float lo, float hi) { f32x2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
TRT_DEV void upk2(f32x2 v, float* lo, float* hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(*lo), "=f"(*hi) : "l"(v)); }
TRT_DEV f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
TRT_DEV f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
TRT_DEV f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// ------------------------------------------------------------------ AABB slab test, hittable/aabb.rs:36-61
// Returns entry (`start`) and exit (`end`) of the range [t_min, +inf) clipped by the three slabs, updated with the
// reference's exact compare/select sequence (NaN from 0*inf leaves the range untouched; swap only when t1 < t0).
// The reference's early-outs `end <= start` after each axis are equivalent to one test at the end because start only
// grows and end only shrinks; the caller combines it with the current closest hit: pass <=> min(hi, end) > start.
TRT_DEV void slab(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, V3 o, V3 inv, float t_min,
                  float* start_out, float* end_out) {
    float start = t_min, end = TRT_INF;
    {
        float t0 = (mnx - o.x) * inv.x, t1 = (mxx - o.x) * inv.x;
        if (t1 < t0) { float tmp = t0; t0 = t1; t1 = tmp; }
        if (start < t0) start = t0;
        if (t1 < end) end = t1;
    }
    {
        float t0 = (mny - o.y) * inv.y, t1 = (mxy - o.y) * inv.y;
        if (t1 < t0) { float tmp = t0; t0 = t1; t1 = tmp; }
        if (start < t0) start = t0;
        if (t1 < end) end = t1;
    }
    {
        float t0 = (mnz - o.z) * inv.z, t1 = (mxz - o.z) * inv.z;
        if (t1 < t0) { float tmp = t0; t0 = t1; t1 = tmp; }
        if (start < t0) start = t0;
        if (t1 < end) end = t1;
    }
    *start_out = start; *end_out = end;
}
// The same function for rays that cannot produce a NaN in (plane - o) * inv — finite origin, finite non-zero 1/d on all
// three axes (`ray_is_tame`): then t0, t1 are ordered numbers, the conditional swap is min/max and the conditional
// updates are max/min, so FMNMX gives bit-identical start/end with a third of the instructions. Every other ray
// (a zero or denormal direction component, NaN/inf anywhere) takes slab() above.
TRT_DEV void slab_tame(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, V3 o, V3 inv, float t_min,
                       float* start_out, float* end_out) {
    float ax = (mnx - o.x) * inv.x, bx = (mxx - o.x) * inv.x;
    float ay = (mny - o.y) * inv.y, by = (mxy - o.y) * inv.y;
    float az = (mnz - o.z) * inv.z, bz = (mxz - o.z) * inv.z;
    *start_out = fmaxf(fmaxf(t_min, fminf(ax, bx)), fmaxf(fminf(ay, by), fminf(az, bz)));
    *end_out = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}
TRT_DEV bool ray_is_tame(V3 o, V3 inv) {
    // all six finite (a NaN or inf makes the sum NaN or inf) and no 1/d component zero
    float si = (fabsf(inv.x) + fabsf(inv.y)) + fabsf(inv.z), so = (fabsf(o.x) + fabsf(o.y)) + fabsf(o.z);
    float mi = fminf(fminf(fabsf(inv.x), fabsf(inv.y)), fabsf(inv.z));
    return si < TRT_INF && so < TRT_INF && mi > 0.0f;
}
// Both children of one node at once (the node stores (left, right) pairs per coordinate, trt_layout.h): the reference's
// `(b - o) * inv` for the two boxes in one FADD2 + one FMUL2 per plane, then slab_tame's min/max per child. Bit-identical.
TRT_DEV void slab2_tame(float4 n0, float4 n1, float4 n2, V3 o, V3 inv, float t_min, float* sl, float* el, float* sr, float* er) {
    const f32x2 ox = pk2(o.x, o.x), oy = pk2(o.y, o.y), oz = pk2(o.z, o.z);
    const f32x2 ix = pk2(inv.x, inv.x), iy = pk2(inv.y, inv.y), iz = pk2(inv.z, inv.z);
    float axl, axr, bxl, bxr, ayl, ayr, byl, byr, azl, azr, bzl, bzr;
    upk2(mul2(sub2(pk2(n0.x, n0.y), ox), ix), &axl, &axr); upk2(mul2(sub2(pk2(n1.z, n1.w), ox), ix), &bxl, &bxr);
    upk2(mul2(sub2(pk2(n0.z, n0.w), oy), iy), &ayl, &ayr); upk2(mul2(sub2(pk2(n2.x, n2.y), oy), iy), &byl, &byr);
    upk2(mul2(sub2(pk2(n1.x, n1.y), oz), iz), &azl, &azr); upk2(mul2(sub2(pk2(n2.z, n2.w), oz), iz), &bzl, &bzr);
    *sl = fmaxf(fmaxf(t_min, fminf(axl, bxl)), fmaxf(fminf(ayl, byl), fminf(azl, bzl)));
    *el = fminf(fminf(fmaxf(axl, bxl), fmaxf(ayl, byl)), fmaxf(azl, bzl));
    *sr = fmaxf(fmaxf(t_min, fminf(axr, bxr)), fmaxf(fminf(ayr, byr), fminf(azr, bzr)));
    *er = fminf(fminf(fmaxf(axr, bxr), fmaxf(ayr, byr)), fmaxf(azr, bzr));
}
TRT_DEV float rcp_approx(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ------------------------------------------------------------------ primitives
// Sphere::hit, hittable/sphere.rs:29-45. Returns t in [t_min, hi) or NaN-free "miss" as false.
// Three shortcuts skip work whose outcome is already decided; each is EXACT — it returns what the literal sequence would — for
// t_min > 0 (the reference's 0.001, cpu.rs:48) and is switched off otherwise:
//  (1) half_b > 0 and c >= 0 (origin on or outside the sphere, moving away: every ray leaving the sphere it starts on — the
//      commonest sphere test of all, because the origin lies inside that sphere's own box): a*c >= 0, so disc = fl(hb*hb - a*c)
//      <= fl(hb*hb) and sqrt(disc) <= sqrt(fl(hb*hb)) = |hb| (sqrt(fl(x*x)) == |x| in binary round-to-nearest) -> both
//      numerators -hb -+ sq are <= 0 -> both roots <= 0 < t_min: a miss, without the square root and the two divisions.
//  (2) c < 0 (origin inside): a*c <= 0 -> sqrt(disc) >= |hb| -> the first numerator -hb - sq <= 0 -> root 1 <= 0 < t_min: its
//      division is skipped, root 2 is computed as usual.
//  (3) root 1 >= hi: root 2 >= root 1 (same positive divisor, larger numerator, correctly rounded division is monotone) -> miss.
//  A NaN anywhere fails the comparisons of (1) and (2) and takes the literal path.
TRT_DEV bool sphere_hit_rec(const float4 s, V3 o, V3 d, float t_min, float hi, float* t_out) {     // s = (centre, radius), already fetched
    V3 oc = o - mk(s.x, s.y, s.z);
    float a = dot(d, d);
    float half_b = dot(oc, d);
    float c = dot(oc, oc) - s.w * s.w;
    const bool pos = t_min > 0.0f && a > 0.0f;
    if (pos && half_b > 0.0f && c >= 0.0f) return false;                     // (1)
    float disc = half_b * half_b - a * c;
    if (disc < 0.0f) return false;
    float sq = sqrtf(disc);
    float t;
    if (!(pos && c < 0.0f)) {                                                 // (2)
        t = (-half_b - sq) / a;
        if (t_min <= t && t < hi) { *t_out = t; return true; }
        if (pos && t >= hi) return false;                                     // (3)
    }
    t = (-half_b + sq) / a;
    if (!(t_min <= t && t < hi)) return false;
    *t_out = t; return true;
}
template <bool SMEM> TRT_DEV bool sphere_hit(const DevScene& S, uint32_t idx, V3 o, V3 d, float t_min, float hi, float* t_out) {
    return sphere_hit_rec(ld4<SMEM>(S.spheres + idx), o, d, t_min, hi, t_out);
}
// Quad::hit, hittable/quad.rs:33-41
// `hi` is the range end of the reference's `contains`; `hi_cut` (>= any t the caller could still accept) only feeds the pre-reject.
// q0 = the quad's first float4 (n, D), already fetched; the other three are read only when the plane is hit inside the range
template <bool SMEM> TRT_DEV bool quad_hit_rec(const float4* q, const float4 q0, V3 o, V3 d, float t_min, float hi, float hi_cut, float* t_out) {
    V3 n = mk(q0.x, q0.y, q0.z);
    float dir_norm = dot(d, n);
    float num = q0.w - dot(o, n);
    // Cheap conservative reject before the IEEE division: num * rcp.approx(dn) is within 2^-22 relative of the exact
    // quotient, so a value clearly outside [t_min, hi) means the exact t is outside too. This removes the division
    // (and its slow path for num == 0) for the common case of a ray leaving the very quad it starts on.
    float ta = num * rcp_approx(dir_norm);
    if (ta < t_min - fabsf(t_min) * 1e-6f || ta > hi_cut + fabsf(hi_cut) * 1e-6f) return false;
    float t = num / dir_norm;
    if (!(t_min <= t && t < hi)) return false;
    float4 q1 = ld4<SMEM>(q + 1), q2 = ld4<SMEM>(q + 2), q3 = ld4<SMEM>(q + 3);
    V3 p = ray_at(o, d, t) - mk(q1.x, q1.y, q1.z);
    V3 w = mk(q1.w, q2.w, q3.w);
    float px = dot(cross(p, mk(q3.x, q3.y, q3.z)), w);
    float py = dot(cross(mk(q2.x, q2.y, q2.z), p), w);
    if (!(0.0f <= px && px < 1.0f && 0.0f <= py && py < 1.0f)) return false;
    *t_out = t; return true;
}
template <bool SMEM> TRT_DEV bool quad_hit(const DevScene& S, uint32_t idx, V3 o, V3 d, float t_min, float hi, float hi_cut, float* t_out) {
    const float4* q = S.quads + 4 * (size_t)idx;
    return quad_hit_rec<SMEM>(q, ld4<SMEM>(q), o, d, t_min, hi, hi_cut, t_out);
}

// ------------------------------------------------------------------ closest hit = BVH::hit / Node::hit, hittable/bvh.rs:25-27,88-107
// ORDERED = false (TRT_BVH_REFERENCE): the tree has the reference's topology and is walked left-first. A deferred right
//   child is re-checked against the closest hit found in the left subtree when it is popped — exactly the narrowed
//   range `t_range.start..hit_record.t` of bvh.rs:97 — and a primitive is accepted only for t < closest (half-open
//   range), so the leftmost leaf wins ties, as in the reference.
// ORDERED = true (TRT_BVH_SAH): any topology, nearer child first. A subtree is culled only when its entry is beyond
//   the closest hit; a primitive replaces the closest hit if its t is smaller, or equal with a smaller reference leaf
//   rank (the reference's tie rule), so the result does not depend on the visiting order.
//
// SIMT shape: one loop whose body has exactly one interior-node region (A) and one primitive region (B). A lane is in
// A when its current ref is an interior node and in B when it is a leaf (possibly the leaf A just selected), so each
// region's code is issued once per iteration for all lanes that need it instead of once per call site.
struct Hit { float t; uint32_t ref; };
// Ordered kinds cull a subtree whose box entry lies beyond the closest hit. A grazing sphere's computed t can be smaller
// than the computed entry of its own box; for such a primitive the reference's answer depends on its own
// visiting order, and the walk must at least FIND it to notice (needs_replay below). Culling only beyond (1 + 2^-8) x the
// closest hit keeps every primitive whose t/entry disagreement is below 0.4 % in play (a sphere's is bounded by
// ~sqrt(20 * 2^-24) = 0.1 %), at the price of a few extra node visits.
#define TRT_WIDE_SMEM_F4 7u         /* float4 per wide node in SHARED memory (8 in global memory), see stage_scene */
#ifndef TRT_SAH_CULL_MARGIN
#define TRT_SAH_CULL_MARGIN 0.00390625f
#endif
TRT_DEV float cull_limit(float best) { return best + fabsf(best) * TRT_SAH_CULL_MARGIN; }

// The traversal stack: (child ref, box entry) pairs in ONE array of 8-byte entries in local memory, so that a push or a pop is
// one STL.64 / LDL.64 (two 32-bit arrays, round 1: two accesses each; merged: 10^6 spheres +2.9 %, final scene +0.6 %).
// Measured and rejected: the first 12 / 16 / 24 entries of every thread in shared memory ([entry][thread], conflict-free):
// 10^6 spheres -3 / -3 / -8 %, final scene -9 % — the shared memory it takes costs more L1 than the stack traffic did.
struct TravStack {
    uint2 e_[TRT_MAX_STACK];
    TRT_DEV void put(int i, uint32_t ref, float s) { e_[i] = make_uint2(ref, __float_as_uint(s)); }
    TRT_DEV void get(int i, uint32_t* ref, float* s) const { const uint2 v = e_[i]; *ref = v.x; *s = __uint_as_float(v.y); }
};

// Resumable form: trav_begin() sets the state up, every trav_step() runs one iteration of the loop (region A, region B,
// pop). closest_hit() is begin + step until done; the wavefront extend kernel interleaves steps of different rays.
struct Trav {
    V3 o, d, inv; float t_min; bool tame;
    uint32_t cur; int sp; Hit best;
    float cur_s;        // ordered kinds: box entry of the child in `cur` (for a leaf: of the primitive's own box)
    bool best_anom;     // ordered kinds: the closest hit's t is smaller than its own box's entry
    bool replay;        // ordered kinds: this ray is being re-traced on the reference-topology tree (see needs_replay)
    float e8;           // wide-8 walk: the margin E of the conservative quantised-box test for this ray (trav8_setup)
};

template <bool COUNT>
TRT_DEV void trav_begin(Trav& T, const DevScene& S, V3 o, V3 d, float t_min, EventCounters* ec) {
    T.o = o; T.d = d; T.t_min = t_min;
    T.best.t = TRT_INF; T.best.ref = TRT_REF_INVALID; T.sp = 0;
    T.cur_s = 0.0f; T.best_anom = false; T.replay = false;
    T.cur = S.root;
    T.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);           // aabb.rs:42, identical for every node of this ray
    T.tame = ray_is_tame(o, T.inv);
    if (COUNT && T.cur != TRT_REF_INVALID) {
        ec->node_visits += 1;          // the root's own Node::hit call (bvh.rs:25-27); its box test is implied by the
        float s0, e0;                  // children's tests, so it is only evaluated to keep the event counts faithful
        slab(S.root_min[0], S.root_min[1], S.root_min[2], S.root_max[0], S.root_max[1], S.root_max[2], o, T.inv, t_min, &s0, &e0);
        if (!(e0 > s0)) T.cur = TRT_REF_INVALID;
    }
}

template <bool ORDERED, bool SMEM, bool COUNT>
TRT_DEV void trav_step(Trav& T, const DevScene& S, TravStack& K, EventCounters* ec) {
    uint32_t cur = T.cur;
    const V3 o = T.o, d = T.d, inv = T.inv; const float t_min = T.t_min;
    if (!(cur & TRT_REF_LEAF)) {
        // ---- region A: one interior node = two child boxes (aabb.rs:36-61 twice)
        TRT_ASSERT(cur < S.n_nodes);
        const float4* np = S.nodes + 4 * (size_t)cur;
        float4 n0 = ld4<SMEM>(np), n1 = ld4<SMEM>(np + 1), n2 = ld4<SMEM>(np + 2), n3 = ld4<SMEM>(np + 3);
        uint32_t lref = __float_as_uint(n3.x), rref = __float_as_uint(n3.y);
        float sl, el, sr, er;
        if (T.tame) {
            slab2_tame(n0, n1, n2, o, inv, t_min, &sl, &el, &sr, &er);
        } else {
            slab(n0.x, n0.z, n1.x, n1.z, n2.x, n2.z, o, inv, t_min, &sl, &el);
            slab(n0.y, n0.w, n1.y, n1.w, n2.y, n2.w, o, inv, t_min, &sr, &er);
        }
        if (COUNT) ec->node_visits += (rref == TRT_REF_INVALID) ? 0u : 2u;   // left.hit + right.hit
        bool pl, pr;
        if (ORDERED) { const float lim = cull_limit(T.best.t); pl = el > sl && !(lim < sl); pr = er > sr && !(lim < sr); }
        else { pl = el > sl && T.best.t > sl; pr = er > sr; }                // right: range end re-checked when it is entered
        pr = pr && rref != TRT_REF_INVALID;
        bool right_first = ORDERED && pl && pr && sr < sl;
        uint32_t c0 = right_first ? rref : lref, c1 = right_first ? lref : rref;
        float s1 = right_first ? sl : sr;
        bool p0 = right_first ? pr : pl, p1 = right_first ? pl : pr;
        if (p0) {
            cur = c0; if (ORDERED) T.cur_s = right_first ? sr : sl;
            if (p1) { TRT_ASSERT(T.sp < TRT_MAX_STACK); K.put(T.sp, c1, s1); ++T.sp; }
        } else if (p1 && (ORDERED || T.best.t > s1)) { cur = c1; if (ORDERED) T.cur_s = s1; }
        else cur = TRT_REF_INVALID;
    }
    if (cur != TRT_REF_INVALID && (cur & TRT_REF_LEAF)) {
        // ---- region B: one primitive (sphere.rs:29-45 / quad.rs:33-41)
        uint32_t idx = cur & TRT_REF_INDEX_MASK;
        TRT_ASSERT(idx < ((cur & TRT_REF_QUAD) ? S.n_quads : S.n_spheres));
        float t; bool h;
        if (cur & TRT_REF_QUAD) { if (COUNT) ec->quad_tests++; h = quad_hit<SMEM>(S, idx, o, d, t_min, ORDERED ? TRT_INF : T.best.t, T.best.t, &t); }
        else { if (COUNT) ec->sphere_tests++; h = sphere_hit<SMEM>(S, idx, o, d, t_min, ORDERED ? TRT_INF : T.best.t, &t); }
        if (h) {
            if (!ORDERED) { T.best.t = t; T.best.ref = cur; }
            else if (t < T.best.t) { T.best.t = t; T.best.ref = cur; T.best_anom = T.cur_s > t; }
            else if (t == T.best.t) {      // exact tie: the lower reference leaf rank wins (both ranks are fetched; ties are rare)
                const uint32_t rb = (T.best.ref & TRT_REF_QUAD) ? ld4ua(S.quad_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y
                                                                : ld4ua(S.sphere_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y;
                const uint32_t rk = (cur & TRT_REF_QUAD) ? ld4ua(S.quad_meta + idx).y : ld4ua(S.sphere_meta + idx).y;
                if (rk < rb) { T.best.ref = cur; T.best_anom = T.cur_s > t; }
            }
        }
        cur = TRT_REF_INVALID;
    }
    // ---- pop: skip entries that the closest hit found meanwhile has cut off
    while (cur == TRT_REF_INVALID && T.sp > 0) {
        --T.sp;
        uint32_t r; float s;
        K.get(T.sp, &r, &s);
        if (ORDERED ? !(cull_limit(T.best.t) < s) : (T.best.t > s)) { cur = r; if (ORDERED) T.cur_s = s; }
    }
    T.cur = cur;
}

// The ordered walk returns the (t, rank)-smallest valid primitive W. That is the reference's answer whenever W's own box
// entry does not exceed its t (nothing valid is closer, so the reference's scan cannot have narrowed its range below W's
// box before reaching it). When it does exceed it — fp cancellation on a grazing hit, a few rays in 10^6 on the sphere
// scenes — the reference's answer depends on its own visiting order, and the ray is replayed on the REFERENCE-topology
// tree (wide nodes, global memory), which reproduces that order. Exact for every hit whose t/entry disagreement is below
// the culling margin above (spheres: always). Every ordered kind carries a reference tree: the host builds it for TRT_BVH_SAH,
// the device's own Node::new (trt_lbvh.cu k_ref_*) for TRT_BVH_LBVH_DEVICE and for big SAH scenes.
template <bool ORDERED, bool SMEM, bool COUNT> TRT_DEV Hit closest_hit4(const DevScene& S, V3 o, V3 d, float t_min, EventCounters* ec);
TRT_DEV bool needs_replay(const Trav& T, const DevScene& S) { return T.best_anom && !T.replay && T.best.ref != TRT_REF_INVALID && S.nodes_ref4 != nullptr; }
TRT_DEV Hit ordered_replay(const DevScene& S, V3 o, V3 d, float t_min) {
    DevScene R = S;
    R.nodes4 = S.nodes_ref4; R.n_nodes4 = S.n_nodes_ref4; R.root = S.root_ref4; R.spheres = S.g_spheres; R.quads = S.g_quads;
    return closest_hit4<false, false, false>(R, o, d, t_min, nullptr);
}

template <bool ORDERED, bool SMEM, bool COUNT>
TRT_DEV Hit closest_hit(const DevScene& S, V3 o, V3 d, float t_min, EventCounters* ec) {
    Trav T;
    trav_begin<COUNT>(T, S, o, d, t_min, ec);
    TravStack K;
    while (T.cur != TRT_REF_INVALID) trav_step<ORDERED, SMEM, COUNT>(T, S, K, ec);
    if (ORDERED && needs_replay(T, S)) return ordered_replay(S, o, d, t_min);   // also when counting: the replay's own events are not counted, the RESULT must not depend on the flag
    return T.best;
}

// ------------------------------------------------------------------ traversal on 4-wide nodes
// A wide node's children are the grandchildren of a binary node in left-to-right order (trt_scene.cpp collapse4).
// ORDERED (TRT_BVH_SAH, device-built tree): same rules as trav_step<true>: a child is entered when its box is crossed and
//   its entry is not beyond the closest hit, a primitive wins with a smaller t or an equal t and a smaller reference leaf
//   rank, so the result does not depend on the visiting order. The crossed children are sorted by entry distance; the
//   nearest is entered, the others are pushed farthest first.
// !ORDERED (TRT_BVH_REFERENCE): children are taken strictly left to right — the first now, the others pushed and
//   re-checked against the narrowed closest hit when popped — so the leaves are met in the reference's order, each with
//   its own box tested against the range the reference would use (lemma L: the skipped intermediate box cannot cull).
// Either way only the number of dependent node fetches per ray changes (about half of the binary walk).
// COUNT (TRT_FLAG_COUNT_WALKED): ec->node_visits counts the wide records fetched (7 x 16 B each), sphere_tests / quad_tests the
// primitive tests — the traffic of THIS walk, for the memory roofline of big scenes (bench.py); results are unaffected.
#ifndef TRT_NODE_V256
#define TRT_NODE_V256 true     /* wide nodes in global memory are fetched with 256-bit loads (see ldg256) */
#endif
// Pop: stack entries that the closest hit found meanwhile has cut off are skipped. With the stack in local memory every entry
// is a DEPENDENT load, and on 10^6 spheres the one-at-a-time loop ran 2.8 times per node step with 2 of 32 lanes active — 31 % of
// the kernel's long-scoreboard stalls (profiles/r2h_tree_c4_2spp): BATCH entries are therefore fetched together and the topmost
// live one is taken. Same entries skipped, same entry taken: results unchanged.
#ifndef TRT_POP_BATCH
#define TRT_POP_BATCH 4      /* 10^6 spheres: 1 -> 2 479, 2 -> 2 482, 4 -> 2 498, 8 -> 2 424 Mrays/s: within noise of each other; 4 kept */
#endif
template <bool ORDERED, int BATCH>
TRT_DEV uint32_t trav_pop(Trav& T, const TravStack& K) {
    uint32_t cur = TRT_REF_INVALID;
    while (cur == TRT_REF_INVALID && T.sp > 0) {
        if (BATCH <= 1) {
            --T.sp;
            uint32_t r; float s;
            K.get(T.sp, &r, &s);
            if (ORDERED ? !(cull_limit(T.best.t) < s) : (T.best.t > s)) { cur = r; if (ORDERED) T.cur_s = s; }
        } else {
            const float lim = ORDERED ? cull_limit(T.best.t) : T.best.t;
            uint32_t r[BATCH > 1 ? BATCH : 1]; float s[BATCH > 1 ? BATCH : 1];
#pragma unroll
            for (int j = 0; j < BATCH; ++j) { const int i = T.sp - 1 - j; K.get(i > 0 ? i : 0, &r[j], &s[j]); }
            int used = T.sp < BATCH ? T.sp : BATCH;          // all of them cut off: drop them all
#pragma unroll
            for (int j = BATCH - 1; j >= 0; --j)
                if (j < T.sp && (ORDERED ? !(lim < s[j]) : (lim > s[j]))) { cur = r[j]; if (ORDERED) T.cur_s = s[j]; used = j + 1; }
            T.sp -= used;
        }
    }
    return cur;
}
// Four child boxes of a wide node STAGED IN SHARED MEMORY, for a tame ray: instead of computing both crossings of an axis and
// ordering them with min / max (slab2_tame: 12 FMNMX per pair of boxes), the plane pair the ray meets FIRST and the one it meets
// LAST are fetched through sign-dependent addresses (1/d > 0: the min planes first) — per axis a tame ray's crossings are ordered
// by the sign alone ((mn - o) <= (mx - o) and a multiplication by a positive / negative number keeps / reverses the order, both
// under round-to-nearest), so start and end are the SAME values slab2_tame computes, bit for bit; an absent child (box
// +inf..-inf) comes out as start = +inf, end = -inf instead of a crossed box and is dropped either way (its ref is INVALID).
// 12 LDS.64 + 6 address adds instead of 6 LDS.128, 24 FMNMX fewer per node. Record layout (trt_layout.h): pair group = 48
// bytes: min x @0, min y @8, min z @16, max x @24, max y @32, max z @40; second group at +48.
#ifndef TRT_SMEM_SIGN_SELECT
#define TRT_SMEM_SIGN_SELECT 1      /* final scene (B200): 8 588 -> 9 054 Mrays/s at 64 spp, 128.0 -> 120.6 ms at 500 spp; reference-topology tree 6 713 -> 7 023 (profiles/r2k_probe_c2_sign_select.log) */
#endif
TRT_DEV f32x2 lds64(uint32_t addr) { f32x2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr)); return v; }
TRT_DEV void slab4_tame_signed(const float4* np, V3 o, V3 inv, float t_min, float* s0, float* e0, float* s1, float* e1,
                               float* s2, float* e2, float* s3, float* e3) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(np);
    const uint32_t kx = inv.x < 0.0f ? 24u : 0u, ky = inv.y < 0.0f ? 24u : 0u, kz = inv.z < 0.0f ? 24u : 0u;
    const uint32_t nx = base + kx, fx = base + 24u - kx, ny = base + 8u + ky, fy = base + 32u - ky, nz = base + 16u + kz, fz = base + 40u - kz;
    const f32x2 ox = pk2(o.x, o.x), oy = pk2(o.y, o.y), oz = pk2(o.z, o.z);
    const f32x2 ix = pk2(inv.x, inv.x), iy = pk2(inv.y, inv.y), iz = pk2(inv.z, inv.z);
    float nxl, nxr, fxl, fxr, nyl, nyr, fyl, fyr, nzl, nzr, fzl, fzr;
    upk2(mul2(sub2(lds64(nx), ox), ix), &nxl, &nxr); upk2(mul2(sub2(lds64(fx), ox), ix), &fxl, &fxr);
    upk2(mul2(sub2(lds64(ny), oy), iy), &nyl, &nyr); upk2(mul2(sub2(lds64(fy), oy), iy), &fyl, &fyr);
    upk2(mul2(sub2(lds64(nz), oz), iz), &nzl, &nzr); upk2(mul2(sub2(lds64(fz), oz), iz), &fzl, &fzr);
    *s0 = fmaxf(fmaxf(t_min, nxl), fmaxf(nyl, nzl)); *e0 = fminf(fminf(fxl, fyl), fzl);
    *s1 = fmaxf(fmaxf(t_min, nxr), fmaxf(nyr, nzr)); *e1 = fminf(fminf(fxr, fyr), fzr);
    upk2(mul2(sub2(lds64(nx + 48u), ox), ix), &nxl, &nxr); upk2(mul2(sub2(lds64(fx + 48u), ox), ix), &fxl, &fxr);
    upk2(mul2(sub2(lds64(ny + 48u), oy), iy), &nyl, &nyr); upk2(mul2(sub2(lds64(fy + 48u), oy), iy), &fyl, &fyr);
    upk2(mul2(sub2(lds64(nz + 48u), oz), iz), &nzl, &nzr); upk2(mul2(sub2(lds64(fz + 48u), oz), iz), &fzl, &fzr);
    *s2 = fmaxf(fmaxf(t_min, nxl), fmaxf(nyl, nzl)); *e2 = fminf(fminf(fxl, fyl), fzl);
    *s3 = fmaxf(fmaxf(t_min, nxr), fmaxf(nyr, nzr)); *e3 = fminf(fminf(fxr, fyr), fzr);
}
template <bool SMEM, bool V256>
TRT_DEV void load_wide_node(const float4* np, float4& a0, float4& a1, float4& a2, float4& b0, float4& b1, float4& b2, float4& rf) {
    if (!SMEM && V256) {
        const F4x2 p0 = ldg256(np), p1 = ldg256(np + 2), p2 = ldg256(np + 4);
        a0 = p0.a; a1 = p0.b; a2 = p1.a; b0 = p1.b; b1 = p2.a; b2 = p2.b; rf = ldg128n(np + 6);
    } else {
        a0 = ld4<SMEM>(np); a1 = ld4<SMEM>(np + 1); a2 = ld4<SMEM>(np + 2);
        b0 = ld4<SMEM>(np + 3); b1 = ld4<SMEM>(np + 4); b2 = ld4<SMEM>(np + 5); rf = ld4<SMEM>(np + 6);
    }
}
template <bool ORDERED, bool SMEM, bool COUNT = false, bool V256 = TRT_NODE_V256>
TRT_DEV void trav4_step(Trav& T, const DevScene& S, float t_min, TravStack& K, EventCounters* ec = nullptr) {   // t_min: the caller's (uniform) value, not a per-lane copy
    uint32_t cur = T.cur;
    const V3 o = T.o, d = T.d, inv = T.inv;
    if (!(cur & TRT_REF_LEAF)) {
        TRT_ASSERT(cur < S.n_nodes4);
        if (COUNT) ec->node_visits += 1;
        const float4* np = S.nodes4 + (size_t)(SMEM ? TRT_WIDE_SMEM_F4 : 8) * cur;
        float4 a0, a1, a2, b0, b1, b2, rf;
        float s0, e0, s1, e1, s2, e2, s3, e3;
        if (SMEM && TRT_SMEM_SIGN_SELECT && T.tame) {
            slab4_tame_signed(np, o, inv, t_min, &s0, &e0, &s1, &e1, &s2, &e2, &s3, &e3);
            rf = ld4<true>(np + 6);
        } else {
            load_wide_node<SMEM, V256>(np, a0, a1, a2, b0, b1, b2, rf);
            if (T.tame) {
                slab2_tame(a0, a1, a2, o, inv, t_min, &s0, &e0, &s1, &e1);
                slab2_tame(b0, b1, b2, o, inv, t_min, &s2, &e2, &s3, &e3);
            } else {
                slab(a0.x, a0.z, a1.x, a1.z, a2.x, a2.z, o, inv, t_min, &s0, &e0);
                slab(a0.y, a0.w, a1.y, a1.w, a2.y, a2.w, o, inv, t_min, &s1, &e1);
                slab(b0.x, b0.z, b1.x, b1.z, b2.x, b2.z, o, inv, t_min, &s2, &e2);
                slab(b0.y, b0.w, b1.y, b1.w, b2.y, b2.w, o, inv, t_min, &s3, &e3);
            }
        }
        uint32_t r0 = __float_as_uint(rf.x), r1 = __float_as_uint(rf.y), r2 = __float_as_uint(rf.z), r3 = __float_as_uint(rf.w);
        const float best = ORDERED ? cull_limit(T.best.t) : T.best.t;
        TRT_ASSERT(T.sp + 3 <= TRT_MAX_STACK);
        if (ORDERED) {
            // a child that is not entered sorts last (key +inf) and carries TRT_REF_INVALID
            if (!(e0 > s0 && !(best < s0) && r0 != TRT_REF_INVALID)) { s0 = TRT_INF; r0 = TRT_REF_INVALID; }
            if (!(e1 > s1 && !(best < s1) && r1 != TRT_REF_INVALID)) { s1 = TRT_INF; r1 = TRT_REF_INVALID; }
            if (!(e2 > s2 && !(best < s2) && r2 != TRT_REF_INVALID)) { s2 = TRT_INF; r2 = TRT_REF_INVALID; }
            if (!(e3 > s3 && !(best < s3) && r3 != TRT_REF_INVALID)) { s3 = TRT_INF; r3 = TRT_REF_INVALID; }
#define TRT_CSWAP(sa, ra, sb, rb) { const bool w_ = sb < sa; const float ts_ = w_ ? sb : sa, tb_ = w_ ? sa : sb; \
                                    const uint32_t tr_ = w_ ? rb : ra, tq_ = w_ ? ra : rb; sa = ts_; sb = tb_; ra = tr_; rb = tq_; }
            TRT_CSWAP(s0, r0, s1, r1) TRT_CSWAP(s2, r2, s3, r3) TRT_CSWAP(s0, r0, s2, r2) TRT_CSWAP(s1, r1, s3, r3) TRT_CSWAP(s1, r1, s2, r2)
#undef TRT_CSWAP
            // the ref, not the key, says whether a slot is occupied
            if (r3 != TRT_REF_INVALID) { K.put(T.sp, r3, s3); ++T.sp; }
            if (r2 != TRT_REF_INVALID) { K.put(T.sp, r2, s2); ++T.sp; }
            if (r1 != TRT_REF_INVALID) { K.put(T.sp, r1, s1); ++T.sp; }
            cur = r0; T.cur_s = s0;
        } else {
            // bvh.rs:97: a later sibling is searched with the range end narrowed by everything before it -> re-checked when popped
            if (e3 > s3 && r3 != TRT_REF_INVALID) { K.put(T.sp, r3, s3); ++T.sp; }
            if (e2 > s2 && r2 != TRT_REF_INVALID) { K.put(T.sp, r2, s2); ++T.sp; }
            if (e1 > s1 && r1 != TRT_REF_INVALID) { K.put(T.sp, r1, s1); ++T.sp; }
            cur = (e0 > s0 && best > s0) ? r0 : TRT_REF_INVALID;
        }
    }
    if (cur != TRT_REF_INVALID && (cur & TRT_REF_LEAF)) {
        uint32_t idx = cur & TRT_REF_INDEX_MASK;
        TRT_ASSERT(idx < ((cur & TRT_REF_QUAD) ? S.n_quads : S.n_spheres));
        float t; bool h;
        if (COUNT) { if (cur & TRT_REF_QUAD) ec->quad_tests += 1; else ec->sphere_tests += 1; }
        if (cur & TRT_REF_QUAD) h = quad_hit<SMEM>(S, idx, o, d, t_min, ORDERED ? TRT_INF : T.best.t, T.best.t, &t);
        else h = sphere_hit<SMEM>(S, idx, o, d, t_min, ORDERED ? TRT_INF : T.best.t, &t);
        if (h) {
            if (!ORDERED) { T.best.t = t; T.best.ref = cur; }
            else if (t < T.best.t) { T.best.t = t; T.best.ref = cur; T.best_anom = T.cur_s > t; }
            else if (t == T.best.t) {      // exact tie: the lower reference leaf rank wins (both ranks are fetched; ties are rare)
                const uint32_t rb = (T.best.ref & TRT_REF_QUAD) ? ld4ua(S.quad_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y
                                                                : ld4ua(S.sphere_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y;
                const uint32_t rk = (cur & TRT_REF_QUAD) ? ld4ua(S.quad_meta + idx).y : ld4ua(S.sphere_meta + idx).y;
                if (rk < rb) { T.best.ref = cur; T.best_anom = T.cur_s > t; }
            }
        }
        cur = TRT_REF_INVALID;
    }
    if (cur == TRT_REF_INVALID) cur = trav_pop<ORDERED, SMEM ? 1 : TRT_POP_BATCH>(T, K);
    T.cur = cur;
}
// One step of the wide walk, or — for a lane that is replaying its ray (needs_replay) — of the reference-order walk on the
// reference tree's wide nodes in global memory. Sharing the loop and the stack keeps the rare replay out of the hot path's
// register budget.
template <bool ORDERED, bool SMEM, bool COUNT = false, bool V256 = TRT_NODE_V256>
TRT_DEV void trav4_any_step(Trav& T, const DevScene& S, float t_min, TravStack& K, EventCounters* ec = nullptr) {
    if (ORDERED && T.replay) {
        DevScene R = S;
        R.nodes4 = S.nodes_ref4; R.n_nodes4 = S.n_nodes_ref4; R.spheres = S.g_spheres; R.quads = S.g_quads;
        trav4_step<false, false, COUNT, V256>(T, R, t_min, K, ec);
    } else trav4_step<ORDERED, SMEM, COUNT, V256>(T, S, t_min, K, ec);
}
// ------------------------------------------------------------------ the same walk, ONE dependent fetch per step (round 2; trees in global memory)
// trav4_step runs region A, then region B for the lanes whose child just turned out to be a leaf, then the pop: a step is a
// chain of up to three dependent fetches (node record, primitive record of the leaf that node selected, stack entry). Here a
// lane is a node OR a leaf for the whole step, decided by the ref it holds when the step begins, and the first 32 bytes of
// either record are fetched through ONE per-lane address before any of the two regions runs; a leaf selected in region A is
// tested in the NEXT step, while the other lanes fetch their next node: one L2 round trip per step instead of up to two, at the
// price of one more step per leaf test. Same tests on the same records in the same per-lane order as trav4_step: results
// are bit-identical. Measured on 10^6 spheres: +-0 while the L1 data pipe was the limiter (see ldg256); kept as
// TRT_FLAG_STEP(3) for A/B.
template <bool ORDERED, bool COUNT>
TRT_DEV void trav4_step_pipe(Trav& T, const DevScene& S, float t_min, TravStack& K, EventCounters* ec = nullptr) {
    uint32_t cur = T.cur;                                   // never TRT_REF_INVALID here
    const V3 o = T.o, d = T.d, inv = T.inv;
    const bool leaf = (cur & TRT_REF_LEAF) != 0u;
    const uint32_t idx = cur & TRT_REF_INDEX_MASK;
    // spheres are 16-byte records: the aligned PAIR that holds sphere idx is fetched (every scene array is padded to 256 bytes)
    const float4* p0 = !leaf ? S.nodes4 + 8 * (size_t)cur : ((cur & TRT_REF_QUAD) ? S.quads + 4 * (size_t)idx : S.spheres + (idx & ~1u));
    const F4x2 f0 = ldg256(p0);                             // issued for node and leaf lanes alike, before either region runs
    if (!leaf) {
        TRT_ASSERT(cur < S.n_nodes4);
        if (COUNT) ec->node_visits += 1;
        const F4x2 f1 = ldg256(p0 + 2), f2 = ldg256(p0 + 4);
        const float4 a0 = f0.a, a1 = f0.b, a2 = f1.a, b0 = f1.b, b1 = f2.a, b2 = f2.b, rf = __ldg(p0 + 6);
        float s0, e0, s1, e1, s2, e2, s3, e3;
        if (T.tame) {
            slab2_tame(a0, a1, a2, o, inv, t_min, &s0, &e0, &s1, &e1);
            slab2_tame(b0, b1, b2, o, inv, t_min, &s2, &e2, &s3, &e3);
        } else {
            slab(a0.x, a0.z, a1.x, a1.z, a2.x, a2.z, o, inv, t_min, &s0, &e0);
            slab(a0.y, a0.w, a1.y, a1.w, a2.y, a2.w, o, inv, t_min, &s1, &e1);
            slab(b0.x, b0.z, b1.x, b1.z, b2.x, b2.z, o, inv, t_min, &s2, &e2);
            slab(b0.y, b0.w, b1.y, b1.w, b2.y, b2.w, o, inv, t_min, &s3, &e3);
        }
        uint32_t r0 = __float_as_uint(rf.x), r1 = __float_as_uint(rf.y), r2 = __float_as_uint(rf.z), r3 = __float_as_uint(rf.w);
        const float best = ORDERED ? cull_limit(T.best.t) : T.best.t;
        TRT_ASSERT(T.sp + 3 <= TRT_MAX_STACK);
        if (ORDERED) {
            if (!(e0 > s0 && !(best < s0) && r0 != TRT_REF_INVALID)) { s0 = TRT_INF; r0 = TRT_REF_INVALID; }
            if (!(e1 > s1 && !(best < s1) && r1 != TRT_REF_INVALID)) { s1 = TRT_INF; r1 = TRT_REF_INVALID; }
            if (!(e2 > s2 && !(best < s2) && r2 != TRT_REF_INVALID)) { s2 = TRT_INF; r2 = TRT_REF_INVALID; }
            if (!(e3 > s3 && !(best < s3) && r3 != TRT_REF_INVALID)) { s3 = TRT_INF; r3 = TRT_REF_INVALID; }
#define TRT_CSWAP(sa, ra, sb, rb) { const bool w_ = sb < sa; const float ts_ = w_ ? sb : sa, tb_ = w_ ? sa : sb; \
                                    const uint32_t tr_ = w_ ? rb : ra, tq_ = w_ ? ra : rb; sa = ts_; sb = tb_; ra = tr_; rb = tq_; }
            TRT_CSWAP(s0, r0, s1, r1) TRT_CSWAP(s2, r2, s3, r3) TRT_CSWAP(s0, r0, s2, r2) TRT_CSWAP(s1, r1, s3, r3) TRT_CSWAP(s1, r1, s2, r2)
#undef TRT_CSWAP
            if (r3 != TRT_REF_INVALID) { K.put(T.sp, r3, s3); ++T.sp; }
            if (r2 != TRT_REF_INVALID) { K.put(T.sp, r2, s2); ++T.sp; }
            if (r1 != TRT_REF_INVALID) { K.put(T.sp, r1, s1); ++T.sp; }
            cur = r0; T.cur_s = s0;
        } else {
            if (e3 > s3 && r3 != TRT_REF_INVALID) { K.put(T.sp, r3, s3); ++T.sp; }
            if (e2 > s2 && r2 != TRT_REF_INVALID) { K.put(T.sp, r2, s2); ++T.sp; }
            if (e1 > s1 && r1 != TRT_REF_INVALID) { K.put(T.sp, r1, s1); ++T.sp; }
            cur = (e0 > s0 && best > s0) ? r0 : TRT_REF_INVALID;
        }
    } else {
        TRT_ASSERT(idx < ((cur & TRT_REF_QUAD) ? S.n_quads : S.n_spheres));
        float t; bool h;
        if (COUNT) { if (cur & TRT_REF_QUAD) ec->quad_tests += 1; else ec->sphere_tests += 1; }
        if (cur & TRT_REF_QUAD) h = quad_hit_rec<false>(p0, f0.a, o, d, t_min, ORDERED ? TRT_INF : T.best.t, T.best.t, &t);      // (n, D)
        else h = sphere_hit_rec((idx & 1u) ? f0.b : f0.a, o, d, t_min, ORDERED ? TRT_INF : T.best.t, &t);                        // (centre, r)
        if (h) {
            if (!ORDERED) { T.best.t = t; T.best.ref = cur; }
            else if (t < T.best.t) { T.best.t = t; T.best.ref = cur; T.best_anom = T.cur_s > t; }
            else if (t == T.best.t) {      // exact tie: the lower reference leaf rank wins
                const uint32_t rb = (T.best.ref & TRT_REF_QUAD) ? ld4ua(S.quad_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y
                                                                : ld4ua(S.sphere_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y;
                const uint32_t rk = (cur & TRT_REF_QUAD) ? ld4ua(S.quad_meta + idx).y : ld4ua(S.sphere_meta + idx).y;
                if (rk < rb) { T.best.ref = cur; T.best_anom = T.cur_s > t; }
            }
        }
        cur = TRT_REF_INVALID;
    }
    if (cur == TRT_REF_INVALID) cur = trav_pop<ORDERED, TRT_POP_BATCH>(T, K);
    T.cur = cur;
}
template <bool ORDERED, bool COUNT>
TRT_DEV void trav4_any_step_pipe(Trav& T, const DevScene& S, float t_min, TravStack& K, EventCounters* ec = nullptr) {
    if (ORDERED && T.replay) {          // the rare replay keeps the plain step
        DevScene R = S;
        R.nodes4 = S.nodes_ref4; R.n_nodes4 = S.n_nodes_ref4; R.spheres = S.g_spheres; R.quads = S.g_quads;
        trav4_step<false, false, COUNT, true>(T, R, t_min, K, ec);
    } else trav4_step_pipe<ORDERED, COUNT>(T, S, t_min, K, ec);
}
// restart the ray of a finished ordered walk in replay mode
TRT_DEV void begin_replay(Trav& T, const DevScene& S) {
    T.replay = true; T.best_anom = false; T.cur = S.root_ref4; T.sp = 0;
    T.best.t = TRT_INF; T.best.ref = TRT_REF_INVALID;
}
// ------------------------------------------------------------------ traversal on compressed 8-wide nodes (round 2; ordered kinds in global memory)
// Node format: trt_lbvh.cu k_cw8_level (96 bytes = three 256-bit loads for up to eight children, against 112 bytes = four loads
// for four). The quantised child boxes only CULL: child k is skipped when the ray provably misses its quantised box, or provably
// enters it beyond the closest hit; everything that decides the RESULT is unchanged — a leaf's own padded box is rebuilt from
// the primitive with the host's operations (sphere, trt_scene.cpp add_sphere) or fetched (quad) and tested with the reference's
// arithmetic, then the primitive, with trav4_step's acceptance and tie rules. The ordered walk's result does not depend on
// the visiting order or on how much is culled, as long as nothing is culled that the exact walk would have to visit:
//   per axis t = fma(q, 2^e * inv, (origin - o) * inv) is within 5 u (B + |o|) |inv| of the real crossing of the quantised
//   plane (u = 2^-24, B = bound8), the reference's (b - o) * inv within 2 u (...) of the real crossing of the exact plane, and
//   the quantised box contains the exact one, so: exact test passes (end > start) => far_q + 14 u X > near_q, and the exact
//   entry s of ANY box inside child k satisfies s >= near_q - 7 u X, with X = max_axis (B + |o|) |inv|. The test uses the margin
//   E = 2^-18 X = 64 u X (it also covers the 3 low bits of a stack entry, used for the slot number while the nearest is picked).
// Only for rays without overflow or NaN (ray_is_flat_fast: |o| < 2^30, |1/d| < 2^64, no zero 1/d; scenes with B <= 2^30);
// every other ray — and a ray whose stack would overflow — is traced in the reference's own order on the reference tree
// (begin_replay), which is exact by construction. The nearest accepted child is entered; the others are pushed as they come.
#define TRT_W8_EPS 3.8146973e-6f      /* 2^-18 */
TRT_DEV bool ray_is_flat_fast(V3 o, V3 inv);
TRT_DEV void trav8_setup(Trav& T, const DevScene& S) {          // right after trav_begin
    if (T.cur == TRT_REF_INVALID) return;
    const V3 o = T.o, inv = T.inv;
    if (!(ray_is_flat_fast(o, inv) && S.bound8 >= 0.0f && T.t_min >= 0.0f)) { begin_replay(T, S); return; }
    const float ex = (S.bound8 + fabsf(o.x)) * fabsf(inv.x), ey = (S.bound8 + fabsf(o.y)) * fabsf(inv.y), ez = (S.bound8 + fabsf(o.z)) * fabsf(inv.z);
    T.e8 = fmaxf(fmaxf(ex, ey), ez) * TRT_W8_EPS + 1e-30f;
    T.cur = 0u;                                                 // root of the wide-8 tree
}
TRT_DEV uint32_t pick8(uint32_t slot, const float4 lo, const float4 hi) {
    const float a = (slot & 2u) ? ((slot & 1u) ? lo.w : lo.z) : ((slot & 1u) ? lo.y : lo.x);
    const float b = (slot & 2u) ? ((slot & 1u) ? hi.w : hi.z) : ((slot & 1u) ? hi.y : hi.x);
    return __float_as_uint((slot & 4u) ? b : a);
}
template <bool COUNT>
TRT_DEV void trav8_step(Trav& T, const DevScene& S, float t_min, TravStack& K, EventCounters* ec = nullptr) {
    uint32_t cur = T.cur;
    const V3 o = T.o, d = T.d, inv = T.inv;
    if (!(cur & TRT_REF_LEAF)) {
        TRT_ASSERT(cur < S.n_nodes8);
        if (T.sp + 7 > TRT_MAX_STACK) { begin_replay(T, S); return; }      // cannot happen on sane trees; exactness must not depend on it
        if (COUNT) ec->node_visits += 1;
        const float4* np = S.nodes8 + 6 * (size_t)cur;
        const F4x2 h0 = ldg256(np), h1 = ldg256(np + 2), h2 = ldg256(np + 4);
        const uint32_t eb = __float_as_uint(h0.a.w);
        const float sx = __uint_as_float((eb & 0xFFu) << 23) * inv.x, sy = __uint_as_float(((eb >> 8) & 0xFFu) << 23) * inv.y, sz = __uint_as_float(((eb >> 16) & 0xFFu) << 23) * inv.z;
        const float bx = (h0.a.x - o.x) * inv.x, by = (h0.a.y - o.y) * inv.y, bz = (h0.a.z - o.z) * inv.z;
        const f32x2 sx2 = pk2(sx, sx), sy2 = pk2(sy, sy), sz2 = pk2(sz, sz), bx2 = pk2(bx, bx), by2 = pk2(by, by), bz2 = pk2(bz, bz);
        const f32x2 two23 = pk2(8388608.0f, 8388608.0f);
        // per axis: the plane words the ray meets first / last (children 0..3 in word 0, 4..7 in word 1)
        const bool px = inv.x > 0.0f, py = inv.y > 0.0f, pz = inv.z > 0.0f;
        const uint32_t lx0 = __float_as_uint(h0.b.x), lx1 = __float_as_uint(h0.b.y), ly0 = __float_as_uint(h0.b.z), ly1 = __float_as_uint(h0.b.w);
        const uint32_t lz0 = __float_as_uint(h1.a.x), lz1 = __float_as_uint(h1.a.y), ux0 = __float_as_uint(h1.a.z), ux1 = __float_as_uint(h1.a.w);
        const uint32_t uy0 = __float_as_uint(h1.b.x), uy1 = __float_as_uint(h1.b.y), uz0 = __float_as_uint(h1.b.z), uz1 = __float_as_uint(h1.b.w);
        const uint32_t nx0 = px ? lx0 : ux0, nx1 = px ? lx1 : ux1, fx0 = px ? ux0 : lx0, fx1 = px ? ux1 : lx1;
        const uint32_t ny0 = py ? ly0 : uy0, ny1 = py ? ly1 : uy1, fy0 = py ? uy0 : ly0, fy1 = py ? uy1 : ly1;
        const uint32_t nz0 = pz ? lz0 : uz0, nz1 = pz ? lz1 : uz1, fz0 = pz ? uz0 : lz0, fz1 = pz ? uz1 : lz1;
        const float E = T.e8, best = cull_limit(T.best.t);
        uint32_t key[8];
#define TRT_W8_CHILD(k, NX, FX, NY, FY, NZ, FZ, R) { \
            const uint32_t sel_ = 0x7540u + ((k) & 3u); \
            float tnx, tfx, tny, tfy, tnz, tfz; \
            upk2(fma2(sub2(pk2(__uint_as_float(__byte_perm(NX, 0x4B000000u, sel_)), __uint_as_float(__byte_perm(FX, 0x4B000000u, sel_))), two23), sx2, bx2), &tnx, &tfx); \
            upk2(fma2(sub2(pk2(__uint_as_float(__byte_perm(NY, 0x4B000000u, sel_)), __uint_as_float(__byte_perm(FY, 0x4B000000u, sel_))), two23), sy2, by2), &tny, &tfy); \
            upk2(fma2(sub2(pk2(__uint_as_float(__byte_perm(NZ, 0x4B000000u, sel_)), __uint_as_float(__byte_perm(FZ, 0x4B000000u, sel_))), two23), sz2, bz2), &tnz, &tfz); \
            const float nr_ = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_min)), fr_ = fminf(fminf(tfx, tfy), tfz); \
            const float sp_ = nr_ - E; \
            const bool ok_ = (fr_ + E > nr_) && !(best < sp_) && __float_as_uint(R) != TRT_REF_INVALID; \
            key[k] = ok_ ? ((__float_as_uint(fmaxf(sp_, 0.0f)) & ~7u) | (uint32_t)(k)) : 0xFFFFFFFFu; }
        TRT_W8_CHILD(0, nx0, fx0, ny0, fy0, nz0, fz0, h2.a.x) TRT_W8_CHILD(1, nx0, fx0, ny0, fy0, nz0, fz0, h2.a.y)
        TRT_W8_CHILD(2, nx0, fx0, ny0, fy0, nz0, fz0, h2.a.z) TRT_W8_CHILD(3, nx0, fx0, ny0, fy0, nz0, fz0, h2.a.w)
        TRT_W8_CHILD(4, nx1, fx1, ny1, fy1, nz1, fz1, h2.b.x) TRT_W8_CHILD(5, nx1, fx1, ny1, fy1, nz1, fz1, h2.b.y)
        TRT_W8_CHILD(6, nx1, fx1, ny1, fy1, nz1, fz1, h2.b.z) TRT_W8_CHILD(7, nx1, fx1, ny1, fy1, nz1, fz1, h2.b.w)
#undef TRT_W8_CHILD
        const uint32_t kmin = min(min(min(key[0], key[1]), min(key[2], key[3])), min(min(key[4], key[5]), min(key[6], key[7])));
        // stack entry: (ref, conservative entry distance = the key without its slot bits: rounded DOWN by < 8 ulp, covered by E)
#define TRT_W8_PUSH(k, R) if (key[k] != 0xFFFFFFFFu && key[k] != kmin) { K.put(T.sp, __float_as_uint(R), __uint_as_float(key[k] & ~7u)); ++T.sp; }
        TRT_W8_PUSH(0, h2.a.x) TRT_W8_PUSH(1, h2.a.y) TRT_W8_PUSH(2, h2.a.z) TRT_W8_PUSH(3, h2.a.w)
        TRT_W8_PUSH(4, h2.b.x) TRT_W8_PUSH(5, h2.b.y) TRT_W8_PUSH(6, h2.b.z) TRT_W8_PUSH(7, h2.b.w)
#undef TRT_W8_PUSH
        cur = kmin != 0xFFFFFFFFu ? pick8(kmin & 7u, h2.a, h2.b) : TRT_REF_INVALID;
    }
    if (cur != TRT_REF_INVALID && (cur & TRT_REF_LEAF)) {
        // the leaf's own padded box, exactly (bvh.rs:89): rebuilt (sphere.rs:17-18 + aabb.rs:13-20) or fetched (quad), then the primitive
        const uint32_t idx = cur & TRT_REF_INDEX_MASK;
        TRT_ASSERT(idx < ((cur & TRT_REF_QUAD) ? S.n_quads : S.n_spheres));
        float mnx, mny, mnz, mxx, mxy, mxz; float4 g0;
        if (cur & TRT_REF_QUAD) {
            if (COUNT) ec->quad_tests += 1;
            const F4x2 bb = ldg256(S.quad_box + 2 * (size_t)idx);
            g0 = __ldg(S.quads + 4 * (size_t)idx);
            mnx = bb.a.x; mny = bb.a.y; mnz = bb.a.z; mxx = bb.b.x; mxy = bb.b.y; mxz = bb.b.z;
        } else {
            if (COUNT) ec->sphere_tests += 1;
            const F4x2 pr = ldg256(S.spheres + (idx & ~1u));          // the aligned pair that holds sphere idx (arrays are padded to 256 bytes)
            g0 = (idx & 1u) ? pr.b : pr.a;
            const float pad = 0.0001f / 2.0f;
            const float lx = g0.x - g0.w, hx = g0.x + g0.w, ly = g0.y - g0.w, hy = g0.y + g0.w, lz = g0.z - g0.w, hz = g0.z + g0.w;
            mnx = fminf(lx, hx) - pad; mxx = fmaxf(lx, hx) + pad; mny = fminf(ly, hy) - pad; mxy = fmaxf(ly, hy) + pad; mnz = fminf(lz, hz) - pad; mxz = fmaxf(lz, hz) + pad;
        }
        float s, e;
        slab_tame(mnx, mny, mnz, mxx, mxy, mxz, o, inv, t_min, &s, &e);      // wide-8 rays are tame (trav8_setup)
        if (e > s && !(cull_limit(T.best.t) < s)) {
            float t; bool h;
            if (cur & TRT_REF_QUAD) h = quad_hit_rec<false>(S.quads + 4 * (size_t)idx, g0, o, d, t_min, TRT_INF, T.best.t, &t);
            else h = sphere_hit_rec(g0, o, d, t_min, TRT_INF, &t);
            if (h) {
                if (t < T.best.t) { T.best.t = t; T.best.ref = cur; T.best_anom = s > t; }
                else if (t == T.best.t) {      // exact tie: the lower reference leaf rank wins
                    const uint32_t rb = (T.best.ref & TRT_REF_QUAD) ? ld4ua(S.quad_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y
                                                                    : ld4ua(S.sphere_meta + (T.best.ref & TRT_REF_INDEX_MASK)).y;
                    const uint32_t rk = (cur & TRT_REF_QUAD) ? ld4ua(S.quad_meta + idx).y : ld4ua(S.sphere_meta + idx).y;
                    if (rk < rb) { T.best.ref = cur; T.best_anom = s > t; }
                }
            }
        }
        cur = TRT_REF_INVALID;
    }
    while (cur == TRT_REF_INVALID && T.sp > 0) {
        --T.sp;
        uint32_t r; float s;
        K.get(T.sp, &r, &s);
        if (!(cull_limit(T.best.t) < s)) cur = r;
    }
    T.cur = cur;
}
template <bool COUNT>
TRT_DEV void trav8_any_step(Trav& T, const DevScene& S, float t_min, TravStack& K, EventCounters* ec = nullptr) {
    if (T.replay) {
        DevScene R = S;
        R.nodes4 = S.nodes_ref4; R.n_nodes4 = S.n_nodes_ref4; R.spheres = S.g_spheres; R.quads = S.g_quads;
        trav4_step<false, false, COUNT, true>(T, R, t_min, K, ec);
    } else trav8_step<COUNT>(T, S, t_min, K, ec);
}

template <bool COUNT>
TRT_DEV Hit closest_hit8(const DevScene& S, V3 o, V3 d, float t_min, EventCounters* ec) {
    Trav T;
    trav_begin<false>(T, S, o, d, t_min, ec);
    trav8_setup(T, S);
    TravStack K;
    for (;;) {
        while (T.cur != TRT_REF_INVALID) trav8_any_step<COUNT>(T, S, t_min, K, ec);
        if (!needs_replay(T, S)) break;
        begin_replay(T, S);
    }
    return T.best;
}
template <bool ORDERED, bool SMEM, bool COUNT>
TRT_DEV Hit closest_hit4(const DevScene& S, V3 o, V3 d, float t_min, EventCounters* ec) {
    Trav T;
    trav_begin<false>(T, S, o, d, t_min, ec);
    TravStack K;
    for (;;) {
        while (T.cur != TRT_REF_INVALID) trav4_any_step<ORDERED, SMEM, COUNT>(T, S, t_min, K, ec);
        if (!(ORDERED && needs_replay(T, S))) break;
        begin_replay(T, S);
    }
    return T.best;
}

// ------------------------------------------------------------------ camera: pointgen.rs:41-43 + camera.rs:58-66 + vec3extend.rs:45-53
// `r` is Philox block 0 of the stream (pixel, sample, bounce 0): (xi1, xi2, a, b); further blocks feed the lens rejection loop.
// Returns the UN-normalised direction; Ray::new's normalisation (ray.rs:12-14) is applied by the caller.
TRT_DEV void camera_ray(const DevCamera& C, const DevParams& P, uint32_t x, uint32_t y, uint32_t pixel, uint32_t sample, U4 r,
                        V3* o_out, V3* dir_out) {
    float u = ((float)x + u01(r.x)) / (float)(C.width - 1);
    float v = ((float)y + u01(r.y)) / (float)(C.height - 1);
    V3 origin = C.position;
    if (!C.pinhole) {
        float a = u11(r.z), b = u11(r.w);
        uint32_t block = 1;
        while (!(a * a + b * b + 0.0f * 0.0f < 1.0f)) {       // squared_length of (a, b, 0)
            U4 e = philox4x32_10(pixel, sample, 0u, block++, P.seed_lo, P.seed_hi);
            a = u11(e.x); b = u11(e.y);
            if (!(a * a + b * b + 0.0f * 0.0f < 1.0f)) { a = u11(e.z); b = u11(e.w); }
        }
        origin = C.position + a * C.du + b * C.dv;
    }   // else: position + a*0 + b*0 == position bit for bit for every finite (a, b); the stream is counter-based, so
        // not drawing the lens sample shifts nothing else
    *o_out = origin;
    *dir_out = C.upper_left + (u * C.horizontal) - (v * C.vertical) - origin;
}
TRT_DEV void primary_ray(const DevCamera& C, const DevParams& P, uint32_t x, uint32_t y, uint32_t pixel, uint32_t sample, V3* o_out, V3* d_out) {
    U4 r = philox4x32_10(pixel, sample, 0u, 0u, P.seed_lo, P.seed_hi);
    V3 dir;
    camera_ray(C, P, x, y, pixel, sample, r, o_out, &dir);
    *d_out = normalized(dir);
}

// ------------------------------------------------------------------ one hit: HitRecord::new (hittable/mod.rs:27-48) + emitted/scatter (material/*.rs)
struct Surface { V3 point, normal; bool front; uint32_t kind; V3 rgb; float param; };

// The hit record in two halves, so that the megakernel can decide "emit and stop / scatter" from the material alone and
// compute the geometric half once, for the lanes that really scatter:
//   hit_material  material id -> (kind, rgb, param): what Material::emitted / scatter need to know (material/*.rs)
//   hit_geometry  HitRecord::new (hittable/mod.rs:27-48): point = at(t), front_face, unit normal against the ray
// Scenes in global memory pay an L2 round trip for every DEPENDENT load of this block while the whole warp waits, so there the
// two material float4 are fetched together (BOTH) and the primitive's first float4 is fetched before the material is known
// (hit_prim_f4, passed to hit_geometry_f4): meta -> material -> primitive becomes (meta | primitive) -> material.
template <bool SMEM, bool BOTH = !SMEM>
TRT_DEV void hit_material(const DevScene& S, Hit h, Surface* sf) {
    const uint32_t idx = h.ref & TRT_REF_INDEX_MASK;
    const uint32_t mat = (h.ref & TRT_REF_QUAD) ? ld4ua(S.quad_meta + idx).x : ld4ua(S.sphere_meta + idx).x;
    const float4 m0 = ld4a(S.materials + 2 * (size_t)mat);
    if (BOTH) {
        const float4 m1 = ld4a(S.materials + 2 * (size_t)mat + 1);
        sf->kind = __float_as_uint(m0.w);
        sf->rgb = mk(m0.x, m0.y, m0.z);
        sf->param = (sf->kind == TRT_MAT_METAL || sf->kind == TRT_MAT_DIELECTRIC) ? m1.x : 0.0f;
    } else {
        sf->kind = __float_as_uint(m0.w);
        sf->rgb = mk(m0.x, m0.y, m0.z);
        sf->param = (sf->kind == TRT_MAT_METAL || sf->kind == TRT_MAT_DIELECTRIC) ? ld4a(S.materials + 2 * (size_t)mat + 1).x : 0.0f;
    }
}
template <bool SMEM>
TRT_DEV float4 hit_prim_f4(const DevScene& S, Hit h) {       // sphere: (centre, r); quad: (n, D)
    const uint32_t idx = h.ref & TRT_REF_INDEX_MASK;
    return (h.ref & TRT_REF_QUAD) ? ld4<SMEM>(S.quads + 4 * (size_t)idx) : ld4<SMEM>(S.spheres + idx);
}
TRT_DEV void hit_geometry_f4(const DevScene& S, Hit h, V3 o, V3 d, const float4 g0, Surface* sf) {
    const uint32_t idx = h.ref & TRT_REF_INDEX_MASK;
    sf->point = ray_at(o, d, h.t);
    V3 outward, nn;
    if (h.ref & TRT_REF_QUAD) {
        outward = mk(g0.x, g0.y, g0.z);                                       // quad.rs:45: the unnormalised n
        const float4 q = ld4a(S.quad_nn + idx); nn = mk(q.x, q.y, q.z);       // n / |n| never changes: host-side, same ops
    } else {
        outward = sf->point - mk(g0.x, g0.y, g0.z);                           // sphere.rs:46-50
        nn = normalized(outward);
    }
    sf->front = dot(d, outward) < 0.0f;
    sf->normal = sf->front ? nn : -nn;
}
template <bool SMEM>
TRT_DEV void hit_geometry(const DevScene& S, Hit h, V3 o, V3 d, Surface* sf) { hit_geometry_f4(S, h, o, d, hit_prim_f4<SMEM>(S, h), sf); }
template <bool SMEM>
TRT_DEV Surface hit_record(const DevScene& S, Hit h, V3 o, V3 d) {
    Surface sf;
    hit_geometry<SMEM>(S, h, o, d, &sf);
    hit_material<SMEM>(S, h, &sf);
    return sf;
}
// Un-normalised scattered direction for the three scattering materials; `r` is Philox block 0 of (pixel, sample, k).
TRT_DEV V3 scatter_direction(const Surface& sf, V3 d, U4 r) {
    if (sf.kind == TRT_MAT_LAMBERTIAN) {                                      // lambertian.rs:16-22
        V3 uv = normalized(random_in_unit_sphere(u01(r.x), u01(r.y), u01(r.z), false));
        V3 dir = sf.normal + uv;
        if (near_zero(dir)) dir = sf.normal;
        return dir;
    } else if (sf.kind == TRT_MAT_METAL) {                                    // metal.rs:18-25
        V3 reflected = reflect(d, sf.normal);
        return reflected + sf.param * random_in_unit_sphere(u01(r.x), u01(r.y), u01(r.z), true);
    } else {                                                                  // dielectric.rs:26-46
        float ri = sf.front ? 1.0f / sf.param : sf.param;
        float c = fminf(-dot(sf.normal, d), 1.0f);
        float s = sqrtf(1.0f - c * c);
        bool total_reflection = ri * s > 1.0f;
        float sqrt_r0 = (1.0f - ri) / (1.0f + ri);
        float r0 = sqrt_r0 * sqrt_r0;
        float m = 1.0f - c;
        float m2 = m * m;
        float m5 = (m2 * m2) * m;
        float reflectance = r0 + (1.0f - r0) * m5;
        return (total_reflection || reflectance > u01(r.x)) ? reflect(d, sf.normal) : refract(d, sf.normal, ri);
    }
}
// Returns true if the path continues with (o, d) replaced by the scattered ray. `k` is the 1-based hit index (Philox bounce word).
template <bool SMEM>
TRT_DEV bool shade(const DevScene& S, const DevParams& P, Hit h, uint32_t pixel, uint32_t sample, uint32_t k,
                   V3* o, V3* d, V3* atten, V3* color) {
    Surface sf = hit_record<SMEM>(S, h, *o, *d);
    // cpu.rs:49-50: color += cumulated_attenuation * emitted().unwrap_or(0)
    V3 emission = (sf.kind == TRT_MAT_LIGHT) ? sf.rgb : mk(0.0f, 0.0f, 0.0f);
    *color = *color + (*atten) * emission;
    if (sf.kind == TRT_MAT_LIGHT) return false;                               // light.rs:17-19
    U4 r = philox4x32_10(pixel, sample, k, 0u, P.seed_lo, P.seed_hi);
    V3 dir = scatter_direction(sf, *d, r);
    *atten = (*atten) * sf.rgb;                                               // cpu.rs:52
    *o = sf.point; *d = normalized(dir);                                      // Ray::new, ray.rs:12-14
    return true;
}

// ------------------------------------------------------------------ flat list (TRT_BVH_FLAT): closest hit without a tree
// For scenes of at most TRT_FLAT_MAX primitives the tree is replaced by the list of primitives in the reference tree's
// left-to-right leaf order. The reference's Node::hit is then exactly a scan of that list with the running closest hit
// as range end (every ancestor box contains the leaf box, so only the leaf's own box test can cull).
// A SIMT machine runs the box half of that scan with zero control divergence, because every lane tests box i in iteration i:
//   phase A  (uniform loop, box data from the kernel-parameter constant bank): a conservative slab test of every
//            primitive's own padded box against [t_min, inf); the passing ones — a superset of the boxes bvh.rs:89
//            accepts with an open range end — form a 64-bit mask.
//   phase B  (per lane): the primitive test (quad.rs:33-41 / sphere.rs:29-45) of the set bits in rank order, keeping the
//            smallest t (a tie keeps the lower rank): W', the (t, rank)-smallest hit candidate.
//   check    W's own box is tested exactly. If it is crossed, W' is the (t, rank)-smallest VALID primitive W (valid =
//            own box crossed and hit; W is a candidate too and nothing hit is smaller than W'). If moreover W's box
//            entry does not exceed its t — always, except for fp cancellation on grazing hits — the reference's scan
//            returns W whatever it visited before (nothing valid is closer; a tie is lower-ranked). Otherwise, and for
//            rays whose slab arithmetic can overflow or produce NaN, the lane replays the reference's scan literally
//            (flat_scan_reference). No hit candidate at all means no valid primitive: a miss. Exact by construction.
struct FlatPrim { float4 f[TRT_FLAT_F4]; };   // trt_layout.h: f[0..3] geometry; f[4] = (mn.x, mx.x, mn.y, mx.y), f[5] = (mn.z, mx.z, ref bits, 0); f[6] spare
static_assert(sizeof(FlatPrim) == 16 * TRT_FLAT_F4, "FlatPrim must match TRT_FLAT_F4");

// Phase A does not need the exact box test: a SUPERSET of the crossed boxes gives the same winner as long as the winner's
// own box is then tested exactly (the check below), because a primitive that is not the closest valid one never decides
// anything. So phase A runs a cheaper conservative test on the centre/half-width form of the box,
//     m = mid*inv - o*inv,   near = m - half*|inv|,   far = m + half*|inv|      (one FMA each, 9 per box instead of 6+6)
// issued as packed FFMA2 (sm_100 `fma.rn.f32x2`, two boxes per group), and accepts when far + E > near, where
// E = 2.5 * 2^-21 * max_axis (bound + |o|)*|inv| dominates the rounding difference between this form and the reference's
// `(b - o) * inv` (both are within 5.1 * 2^-24 * (bound + |o|)*|inv| of the real value per axis; the host
// rounds `half` up so that [mid - half, mid + half] contains the box). Only taken for rays with |1/d| < 2^64 and
// |o| < 2^30 on every axis and scenes with bound <= 2^30 (no overflow, no NaN); every other ray replays the literal scan.
// Layout: 3 float4 per PAIR of primitives (2j, 2j+1): (m0.x m1.x m0.y m1.y) (m0.z m1.z h0.x h1.x) (h0.y h1.y h0.z h1.z),
// padded with NaN (never accepted) to TRT_FLAT_MAX primitives.
#define TRT_FLAT_CBOX_F4 (3 * TRT_FLAT_MAX / 2)
#define TRT_FLAT_EPS 1.1920929e-6f      /* 2.5 * 2^-21 */
TRT_DEV bool ray_is_flat_fast(V3 o, V3 inv) {
    float si = (fabsf(inv.x) + fabsf(inv.y)) + fabsf(inv.z), so = (fabsf(o.x) + fabsf(o.y)) + fabsf(o.z);   // NaN/inf fail both
    float mi = fminf(fminf(fabsf(inv.x), fabsf(inv.y)), fabsf(inv.z));
    return si < 1.8446744e19f && so < 1.0737418e9f && mi > 0.0f;
}

template <bool DSMEM> TRT_DEV float4 ldf(const float4* p) { return ld4<DSMEM>(p); }

// one primitive against [t_min, hi): returns t or false. hi_cut (>= every t the caller could still accept) only feeds the
// conservative pre-rejects, which skip the IEEE division (quad) or square root and divisions (sphere) when the result is
// clearly out of range; a pre-reject never changes an accepted t.
template <bool DSMEM>
TRT_DEV bool flat_prim_hit(const float4* p, bool is_quad, V3 o, V3 d, float t_min, float hi, float hi_cut, float* t_out) {
    float4 g0 = ldf<DSMEM>(p);
    float t;
    if (is_quad) {
        V3 nrm = mk(g0.x, g0.y, g0.z);
        float dn = dot(d, nrm), num = g0.w - dot(o, nrm);
        float ta = num * rcp_approx(dn);
        if (ta < t_min - fabsf(t_min) * 1e-6f || ta > hi_cut + fabsf(hi_cut) * 1e-6f) return false;
        t = num / dn;
        if (!(t_min <= t && t < hi)) return false;
        float4 g1 = ldf<DSMEM>(p + 1), g2 = ldf<DSMEM>(p + 2), g3 = ldf<DSMEM>(p + 3);
        V3 pp = ray_at(o, d, t) - mk(g1.x, g1.y, g1.z);
        V3 w = mk(g1.w, g2.w, g3.w);
        float px = dot(cross(pp, mk(g3.x, g3.y, g3.z)), w), py = dot(cross(mk(g2.x, g2.y, g2.z), pp), w);
        if (!(0.0f <= px && px < 1.0f && 0.0f <= py && py < 1.0f)) return false;
    } else {
        V3 oc = o - mk(g0.x, g0.y, g0.z);
        float a = dot(d, d), half_b = dot(oc, d), c = dot(oc, oc) - g0.w * g0.w;
        const bool pos = t_min > 0.0f && a > 0.0f;                           // the three exact shortcuts of sphere_hit
        if (pos && half_b > 0.0f && c >= 0.0f) return false;
        float disc = half_b * half_b - a * c;
        if (disc < 0.0f) return false;
        float sq = sqrtf(disc);
        bool have = false;
        if (!(pos && c < 0.0f)) {
            t = (-half_b - sq) / a;
            have = t_min <= t && t < hi;
            if (!have && pos && t >= hi) return false;
        }
        if (!have) { t = (-half_b + sq) / a; if (!(t_min <= t && t < hi)) return false; }
    }
    *t_out = t; return true;
}

// the reference's scan, literally (bvh.rs:88-107 on the leaf list)
template <bool DSMEM>
TRT_DEV Hit flat_scan_reference(const FlatPrim* Dv, uint32_t n, V3 o, V3 d, V3 inv, float t_min) {
    Hit best; best.t = TRT_INF; best.ref = TRT_REF_INVALID;
    for (uint32_t i = 0; i < n; ++i) {
        const float4* p = (const float4*)(Dv + i);
        float4 b0 = ldf<DSMEM>(p + 4), b1 = ldf<DSMEM>(p + 5);
        float s, e;
        slab(b0.x, b0.z, b1.x, b0.y, b0.w, b1.y, o, inv, t_min, &s, &e);
        if (!(e > s && best.t > s)) continue;
        uint32_t ref = __float_as_uint(b1.z);
        float t;
        if (flat_prim_hit<DSMEM>(p, (ref & TRT_REF_QUAD) != 0, o, d, t_min, best.t, best.t, &t)) { best.t = t; best.ref = ref; }
    }
    return best;
}

// CB: the conservative boxes as seen by uniform (same-address) loads — kernel parameters in the megakernel;
// Dv: the primitive list for per-lane indexed loads; qm0 / qm1: bit r = record r is a quad.
template <bool DSMEM>
TRT_DEV Hit closest_hit_flat(const float4* __restrict__ CB, float bound, const FlatPrim* Dv, uint32_t n, uint32_t qm0, uint32_t qm1, V3 o, V3 d, float t_min) {
    Hit best; best.t = TRT_INF; best.ref = TRT_REF_INVALID;
    const V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    if (!(ray_is_flat_fast(o, inv) && bound >= 0.0f)) return flat_scan_reference<DSMEM>(Dv, n, o, d, inv, t_min);
    // ---- phase A: which boxes can the ray cross at all (superset)
    const float aix = fabsf(inv.x), aiy = fabsf(inv.y), aiz = fabsf(inv.z);
    const float E = fmaxf(fmaxf((bound + fabsf(o.x)) * aix, (bound + fabsf(o.y)) * aiy), (bound + fabsf(o.z)) * aiz) * TRT_FLAT_EPS + 1e-30f;
    const float cx = -(o.x * inv.x), cy = -(o.y * inv.y), cz = -(o.z * inv.z);
    // every per-ray operand is a scalar broadcast to both halves (pk2(v, v)); the pairs are the two boxes of a group
    const f32x2 ix2 = pk2(inv.x, inv.x), iy2 = pk2(inv.y, inv.y), iz2 = pk2(inv.z, inv.z);
    const f32x2 cx2 = pk2(cx, cx), cy2 = pk2(cy, cy), cz2 = pk2(cz, cz);
    const f32x2 px2 = pk2(aix, aix), py2 = pk2(aiy, aiy), pz2 = pk2(aiz, aiz);
    const f32x2 nx2 = pk2(-aix, -aix), ny2 = pk2(-aiy, -aiy), nz2 = pk2(-aiz, -aiz);
    uint32_t mask[2] = {0u, 0u};                       // primitives 0..31 and 32..63
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        if (half == 1 && n <= 32u) break;
#pragma unroll
        for (int j = 0; j < 16; ++j) {                 // pair j of this half; static indices -> constant-bank operands
            const int pr = 16 * half + j;
            if ((j & 1) == 0 && 2u * pr >= n) break;   // n is padded to a multiple of 4 with NaN boxes
            const float4 a = CB[3 * pr], b = CB[3 * pr + 1], c = CB[3 * pr + 2];
            const f32x2 hx = pk2(b.z, b.w), hy = pk2(c.x, c.y), hz = pk2(c.z, c.w);
            const f32x2 mx = fma2(pk2(a.x, a.y), ix2, cx2), my = fma2(pk2(a.z, a.w), iy2, cy2), mz = fma2(pk2(b.x, b.y), iz2, cz2);
            float n0x, f0x, n0y, f0y, n0z, f0z, n1x, f1x, n1y, f1y, n1z, f1z;
            upk2(fma2(hx, nx2, mx), &n0x, &n1x); upk2(fma2(hx, px2, mx), &f0x, &f1x);
            upk2(fma2(hy, ny2, my), &n0y, &n1y); upk2(fma2(hy, py2, my), &f0y, &f1y);
            upk2(fma2(hz, nz2, mz), &n0z, &n1z); upk2(fma2(hz, pz2, mz), &f0z, &f1z);
            const float s0 = fmaxf(fmaxf(t_min, n0x), fmaxf(n0y, n0z)), e0 = fminf(fminf(f0x, f0y), f0z);
            const float s1 = fmaxf(fmaxf(t_min, n1x), fmaxf(n1y, n1z)), e1 = fminf(fminf(f1x, f1y), f1z);
            if (e0 + E > s0) mask[half] |= 1u << (2 * j);
            if (e1 + E > s1) mask[half] |= 2u << (2 * j);
        }
    }
    uint32_t best_i = 0xFFFFFFFFu;
    // ---- phase B: primitive tests of the candidates, in rank order, keeping the smallest t (a tie keeps the lower rank).
    // (A two-pass form — cheap can-it-be-hit pass over all candidates, then the exact test of the approximately nearest one first —
    // was measured 19 % SLOWER on Cornell, 18.0 vs 22.2 Grays/s: the second loop's overhead outweighs the exact-test rounds it saves.)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t m = mask[half];
        const uint32_t qm = half ? qm1 : qm0;
        while (m) {
            const uint32_t bit = (uint32_t)__ffs((int)m) - 1u;
            const uint32_t i = bit + 32u * half;
            m &= m - 1u;
            TRT_ASSERT(i < n);
            float t;
            if (flat_prim_hit<DSMEM>((const float4*)(Dv + i), (qm >> bit) & 1u, o, d, t_min, TRT_INF, best.t, &t) && t < best.t) { best.t = t; best_i = i; }
        }
    }
    if (best_i == 0xFFFFFFFFu) return best;
    // ---- check: the winner's own box, exactly (bvh.rs:89), and its entry against its t. (min, max) pairs per axis: the
    // reference's `(b - o) * inv` for both planes of an axis in one FADD2 + one FMUL2 (bit-identical to slab_tame)
    const float4* p = (const float4*)(Dv + best_i);
    const float4 b0 = ldf<DSMEM>(p + 4), b1 = ldf<DSMEM>(p + 5);
    best.ref = __float_as_uint(b1.z);
    float ax, bx, ay, by, az, bz;
    upk2(mul2(sub2(pk2(b0.x, b0.y), pk2(o.x, o.x)), pk2(inv.x, inv.x)), &ax, &bx);
    upk2(mul2(sub2(pk2(b0.z, b0.w), pk2(o.y, o.y)), pk2(inv.y, inv.y)), &ay, &by);
    upk2(mul2(sub2(pk2(b1.x, b1.y), pk2(o.z, o.z)), pk2(inv.z, inv.z)), &az, &bz);
    const float s = fmaxf(fmaxf(t_min, fminf(ax, bx)), fmaxf(fminf(ay, by), fminf(az, bz)));
    const float e = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    if (!(e > s) || s > best.t) return flat_scan_reference<DSMEM>(Dv, n, o, d, inv, t_min);
    return best;
}

// One whole path: CpuSampler::single_point_sampling, renderer/sampler/cpu.rs:39-65
// KIND: 0 = TRT_BVH_REFERENCE, 1 = TRT_BVH_SAH (binary nodes), 2 = TRT_BVH_FLAT, 3 / 4 = SAH / REFERENCE on their 4-wide nodes,
// 5 = ordered tree on its compressed 8-wide nodes (global memory)
template <int KIND, bool SMEM, bool COUNT>
TRT_DEV Hit closest_hit_any(const DevScene& S, V3 o, V3 d, float t_min, EventCounters* ec) {
    if (KIND == 2) return closest_hit_flat<false>(S.flat_cbox, S.flat_bound, S.flat, S.n_flat, S.flat_qmask[0], S.flat_qmask[1], o, d, t_min);
    if (KIND == 5) return closest_hit8<COUNT>(S, o, d, t_min, ec);
    if (KIND == 3 || KIND == 4) return closest_hit4<KIND == 3, SMEM, COUNT>(S, o, d, t_min, ec);
    return closest_hit<KIND == 1, SMEM, COUNT>(S, o, d, t_min, ec);
}

template <int KIND, bool SMEM, bool COUNT>
TRT_DEV V3 trace_path(const DevScene& S, const DevParams& P, V3 o, V3 d, uint32_t pixel, uint32_t sample,
                      uint32_t* n_rays_out, Hit* first_hit, EventCounters* ec) {
    V3 color = mk(0.f, 0.f, 0.f), atten = mk(1.f, 1.f, 1.f);
    uint32_t remain = P.max_bounces, k = 0, rays = 0;
    while (remain > 0) {
        Hit h = closest_hit_any<KIND, SMEM, COUNT>(S, o, d, P.t_min, ec);
        ++rays;
        if (rays == 1 && first_hit) *first_hit = h;
        if (h.ref == TRT_REF_INVALID) { color = color + atten * mk(P.bg[0], P.bg[1], P.bg[2]); break; }   // cpu.rs:58-61
        ++k;
        if (!shade<SMEM>(S, P, h, pixel, sample, k, &o, &d, &atten, &color)) break;
        --remain;
    }
    if (n_rays_out) *n_rays_out = rays;
    return color;
}

// ------------------------------------------------------------------ epilogue arithmetic (shared by k_finalize and the multi-GPU reduce)
// Color::gamma_correction (utils/image.rs:92-98), clamp(0, 0.999) and `* 255 as u8` (utils/image.rs:100-111)
TRT_DEV uint8_t tonemap_u8(float lin) {
    float g = powf(lin, 1.0f / 2.2f);
    float cl = g < 0.0f ? 0.0f : (g > 0.999f ? 0.999f : g);     // NaN stays NaN
    float s = cl * 255.0f;
    return (s == s) ? (uint8_t)s : (uint8_t)0;                   // `as u8`: truncation, NaN -> 0
}

// ------------------------------------------------------------------ shared-memory staging of small scenes
// The traversal data (nodes, primitives, flat list) are staged whenever they fit; the per-hit records (meta, materials,
// quad normals: one access per ray segment) only while they are small — the final scene has one material per sphere, and
// staging its 31 KB of per-hit records would cap the kernel at 3 CTAs per SM instead of 4.
#define TRT_AUX_SMEM_F4 512u       /* 8 KB */
__host__ __device__ inline uint32_t aux_f4(uint32_t n_spheres, uint32_t n_quads, uint32_t n_materials) { return n_spheres + 2u * n_quads + 2u * n_materials; }
// layout: [nodes or wide nodes | spheres | quads | sphere_meta | quad_meta | quad_nn | materials (the four per-hit arrays only while small) | flat list]
template <bool SMEM, int KIND>
__device__ __forceinline__ DevScene stage_scene(const DevScene& G, uint32_t* used_f4 = nullptr) {
    if (used_f4) *used_f4 = 0;
    if (!SMEM) return G;
    extern __shared__ float4 smem_f4[];
    DevScene S = G;
    uint32_t off = 0;
    auto copy = [&](const float4* src, uint32_t n) -> const float4* {
        float4* dst = smem_f4 + off;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
        off += n;
        return dst;
    };
    if (KIND == 3 || KIND == 4) {          // the walk reads only one of the two node arrays
        // wide records are re-strided to 7 float4 (their 8th is unused): an ODD stride puts the same field of different nodes on
        // different bank groups (128-byte records: every lane of a quarter-warp on the same four banks, 53 % of the final scene's
        // shared-memory wavefronts were conflicts, profiles/r2a) and the staged tree is 1/8 smaller
        float4* dst = smem_f4 + off;
        const uint32_t n = TRT_WIDE_SMEM_F4 * G.n_nodes4;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t node = i / TRT_WIDE_SMEM_F4; dst[i] = G.nodes4[8u * node + (i - TRT_WIDE_SMEM_F4 * node)]; }
        off += n; S.nodes4 = dst;
    } else S.nodes = copy(G.nodes, 4 * G.n_nodes);
    S.spheres = copy(G.spheres, G.n_spheres);
    S.quads = copy(G.quads, 4 * G.n_quads);
    if (aux_f4(G.n_spheres, G.n_quads, G.n_materials) <= TRT_AUX_SMEM_F4) {   // else: left in global memory (generic loads)
        S.sphere_meta = (const uint4*)copy((const float4*)G.sphere_meta, G.n_spheres);
        S.quad_meta = (const uint4*)copy((const float4*)G.quad_meta, G.n_quads);
        S.quad_nn = copy(G.quad_nn, G.n_quads);
        S.materials = copy(G.materials, 2 * G.n_materials);
    }
    S.flat = (const FlatPrim*)copy((const float4*)G.flat, TRT_FLAT_F4 * G.n_flat);
    if (used_f4) *used_f4 = off;
    __syncthreads();
    return S;
}

} // namespace trtd
