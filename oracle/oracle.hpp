// oracle/oracle.hpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain C++17 restatement of the CPU path tracer of cheolwanpark/tiny-raytracer, written to be the
// checker for the CUDA path under tiny-raytracer_b200/csrc. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this. It shares NO source with the product:
// every routine here is written from the reference's Rust (cited as `file:line` relative to
// /root/reference/raytracer/src unless stated) and from the numeric contract in DESIGN.md §3.
//
// Parity status ("pinned" = checked against a vector the reference's own tests hold):
//   pinned   : Vec3 length/dot/cross (math/vec3.rs:268-284), Ray::at (ray.rs:33-40), Camera::new basis
//              (camera.rs:78-105), Sphere::hit 3 rays (hittable/sphere.rs:68-100), Quad::hit 3 rays
//              (hittable/quad.rs:67-96), sample bookkeeping (renderer/pointgen.rs:63-107), mean+gamma
//              (renderer/imager.rs:72-108), liveness (renderer/sampler/cpu.rs:131-180), Philox4x32-10
//              Random123 known-answer vectors.
//   UNPINNED : BVH build/traversal, the four materials, reflect/refract, random direction sampling, tone-map
//              quantisation and every rendered image. The reference holds no test vector for them and its RNG is the
//              OS-seeded rand::thread_rng() (utils/random.rs:15-18), so no reference output is reproducible.
//              For these rows this file is the authority: "parity unpinned".
//
// Deliberate, documented deviations from the reference:
//   * RNG: rand 0.8.5 ThreadRng (ChaCha12, unseeded) -> Philox4x32-10 keyed (seed; pixel, sample, bounce, block).
//     u32 -> f32 exactly as rand 0.8.5's UniformFloat<f32>: (bits >> 9) * 2^-23, range(lo,hi) = f*(hi-lo)+lo.
//   * libm sin/cos/acos/cbrt (vec3extend.rs:19-27) -> MATH_DET: functions built only from IEEE + - * / sqrt so the
//     CUDA kernels can reproduce them bit-for-bit. MATH_LIBM keeps the reference's libm call sequence (glibc) and is
//     used to show the two agree statistically.
//   * Lambertian ignores r = cbrt(u3) (the direction is normalised anyway); removes the u3 == 0 NaN of
//     vec3extend.rs:22,32-34.
//   * Accumulation: the reference adds color*(1/spp) in f32 in channel-arrival order (imager.rs:50), which is
//     nondeterministic. Here: f64 sum of the f32 sample colours in sample order, times 1.0/spp, rounded to f32.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <algorithm>
#include <limits>

namespace orc {

typedef float Float;

// ---------------------------------------------------------------- math/vec3.rs
struct Vec3 {
    Float x, y, z;
    Float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline Vec3 v3(Float x, Float y, Float z) { Vec3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline Vec3 diag(Float v) { return v3(v, v, v); }                                   // vec3.rs:24-26
static inline Vec3 operator-(Vec3 a) { return v3(-a.x, -a.y, -a.z); }                      // vec3.rs:62-68
static inline Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); } // vec3.rs:70-76
static inline Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); } // vec3.rs:86-92
static inline Vec3 operator*(Vec3 a, Float s) { return v3(a.x * s, a.y * s, a.z * s); }    // vec3.rs:102-108
static inline Vec3 operator*(Float s, Vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }    // vec3.rs:110-116
static inline Vec3 operator*(Vec3 a, Vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); } // vec3.rs:126-132
static inline Vec3 operator/(Vec3 a, Float s) { return v3(a.x / s, a.y / s, a.z / s); }    // vec3.rs:142-148 (true division)
static inline Float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // vec3.rs:49-51 (left to right)
static inline Vec3 cross(Vec3 a, Vec3 b) {                                                 // vec3.rs:53-59
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline Float squared_length(Vec3 a) { return dot(a, a); }                           // vec3.rs:41-43
static inline Float length(Vec3 a) { return std::sqrt(squared_length(a)); }                // vec3.rs:37-39
static inline Vec3 normalized(Vec3 a) { return a / length(a); }                            // vec3.rs:45-47
static inline bool near_zero(Vec3 a) {                                                     // vec3.rs:32-35
    const Float eps = 1e-7f;
    return std::fabs(a.x) < eps && std::fabs(a.y) < eps && std::fabs(a.z) < eps;
}
// Rust f32::min/max return the non-NaN operand, as fminf/fmaxf do.
static inline Vec3 vmin(Vec3 a, Vec3 b) { return v3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); } // vec3extend.rs:59-65
static inline Vec3 vmax(Vec3 a, Vec3 b) { return v3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); } // vec3extend.rs:67-73
// vec3extend.rs:75-77 : `self - 2.0 * self.dot(normal) * normal` associates as ((2*dot)*n)
static inline Vec3 reflect(Vec3 v, Vec3 n) { return v - (2.0f * dot(v, n)) * n; }
// vec3extend.rs:79-84
static inline Vec3 refract(Vec3 v, Vec3 n, Float etai_over_etat) {
    Float c = fminf(-dot(n, v), 1.0f);
    Vec3 perp = etai_over_etat * (v + n * c);
    Vec3 parallel = (-std::sqrt(std::fabs(1.0f - squared_length(perp)))) * n;
    return parallel + perp;
}

// ---------------------------------------------------------------- ray.rs
struct Ray {
    Vec3 origin, direction;
    Vec3 at(Float t) const { return origin + t * direction; }                              // ray.rs:24-26
};
static inline Ray make_ray(Vec3 o, Vec3 d) { Ray r; r.origin = o; r.direction = normalized(d); return r; } // ray.rs:12-14

struct Range {                       // std::ops::Range<f32>: contains(t) <=> start <= t < end
    Float start, end;
    bool contains(Float t) const { return start <= t && t < end; }
};

// ---------------------------------------------------------------- Philox4x32-10
static inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// rand 0.8.5 UniformFloat<f32>::sample_single: 23 random mantissa bits -> [0,1)
static inline Float u32_to_unit(uint32_t b) { return (Float)(b >> 9) * 1.1920928955078125e-07f; }

// Sequential reader over blocks 0,1,2.. of the stream (seed; pixel, sample, bounce).
struct Stream {
    uint32_t key[2], ctr[4], buf[4]; int pos;
    Stream(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = pixel; ctr[1] = sample; ctr[2] = bounce; ctr[3] = 0; pos = 4;
    }
    uint32_t next_u32() {
        if (pos == 4) { philox4x32_10(ctr, key, buf); ctr[3] += 1; pos = 0; }
        return buf[pos++];
    }
    Float random() { return u32_to_unit(next_u32()); }                                      // utils/random.rs:11-13
    Float random_range(Float lo, Float hi) { return u32_to_unit(next_u32()) * (hi - lo) + lo; } // utils/random.rs:15-18
};

// ---------------------------------------------------------------- deterministic transcendental contract
enum MathMode { MATH_DET = 0, MATH_LIBM = 1 };

// sin and cos of 2*pi*u for u in [0,1), octant reduction on u (exact), cephes-style polynomials on [0, pi/4],
// every operation a separately rounded IEEE op.
static inline void det_sincos2pi(Float u, Float* s_out, Float* c_out) {
    Float q = u * 8.0f;
    int k = (int)q;
    Float f = q - (Float)k;
    Float r = (k & 1) ? (1.0f - f) : f;
    Float x = r * 0.785398163397448309616f;
    Float z = x * x;
    Float sp = -1.9515295891e-4f * z;
    sp = sp + 8.3321608736e-3f;
    sp = sp * z;
    sp = sp - 1.6666654611e-1f;
    sp = sp * z;
    sp = sp * x;
    Float s = sp + x;
    Float cp = 2.443315711809948e-5f * z;
    cp = cp - 1.388731625493765e-3f;
    cp = cp * z;
    cp = cp + 4.166664568298827e-2f;
    cp = cp * z;
    cp = cp * z;
    Float hz = 0.5f * z;
    cp = cp - hz;
    Float c = cp + 1.0f;
    Float ss = s, cc = c;
    if ((k + 1) & 2) { ss = c; cc = s; }
    if (k >= 4) ss = -ss;
    if (k >= 2 && k <= 5) cc = -cc;
    *s_out = ss; *c_out = cc;
}
// cube root for x >= 0: bit-trick seed + 3 Newton steps, IEEE ops only.
static inline Float det_cbrt(Float x) {
    if (!(x > 0.0f)) return 0.0f;
    uint32_t b; std::memcpy(&b, &x, 4);
    b = b / 3u + 709921077u;
    Float y; std::memcpy(&y, &b, 4);
    for (int i = 0; i < 3; ++i) {
        Float y2 = y * y;
        Float y3 = y2 * y;
        Float num = y3 - x;
        Float den = 3.0f * y2;
        y = y - num / den;
    }
    return y;
}

// vec3extend.rs:15-30 new_random_in_unit_sphere. `use_r` = false is the Lambertian fix (r = 1).
static inline Vec3 random_in_unit_sphere(Stream& rng, MathMode mode, bool use_r) {
    Float u1 = rng.random();
    Float u2 = rng.random();
    Float u3 = rng.random();
    Float sin_t, cos_t, sin_phi, cos_phi, r;
    if (mode == MATH_LIBM) {
        Float theta = 2.0f * 3.14159265358979323846f * u1;       // vec3extend.rs:19
        Float phi = std::acos(1.0f - 2.0f * u2);                 // :20
        r = use_r ? std::cbrt(u3) : 1.0f;                        // :21
        sin_phi = std::sin(phi); cos_phi = std::cos(phi);        // :23,27
        sin_t = std::sin(theta); cos_t = std::cos(theta);        // :25-26
    } else {
        det_sincos2pi(u1, &sin_t, &cos_t);
        cos_phi = 1.0f - 2.0f * u2;
        Float s2 = 1.0f - cos_phi * cos_phi;
        sin_phi = std::sqrt(fmaxf(s2, 0.0f));
        r = use_r ? det_cbrt(u3) : 1.0f;
    }
    Float x = r * sin_phi * cos_t;                               // :25
    Float y = r * sin_phi * sin_t;                               // :26
    Float z = r * cos_phi;                                       // :27
    return v3(x, y, z);
}
// vec3extend.rs:32-34
static inline Vec3 random_unit_vector(Stream& rng, MathMode mode) { return normalized(random_in_unit_sphere(rng, mode, false)); }
// vec3extend.rs:45-53
static inline Vec3 random_in_unit_disk(Stream& rng) {
    for (;;) {
        Float a = rng.random_range(-1.0f, 1.0f);
        Float b = rng.random_range(-1.0f, 1.0f);
        Vec3 p = v3(a, b, 0.0f);
        if (squared_length(p) < 1.0f) return p;
    }
}

// ---------------------------------------------------------------- hittable/aabb.rs
struct AABB {
    Vec3 min, max;
    static AABB make(Vec3 a, Vec3 b) {                                                      // aabb.rs:13-20
        const Float PADDING_AMOUNT = 0.0001f;
        Vec3 padding = diag(PADDING_AMOUNT / 2.0f);
        AABB r; r.min = vmin(a, b) - padding; r.max = vmax(a, b) + padding; return r;
    }
    static AABB merge(AABB a, AABB b) { AABB r; r.min = vmin(a.min, b.min); r.max = vmax(a.max, b.max); return r; } // aabb.rs:30-34
    bool intersect(const Ray& ray, Range t_range) const {                                   // aabb.rs:36-61
        Vec3 o = ray.origin, d = ray.direction;
        for (int i = 0; i < 3; ++i) {
            Float mn = min[i], mx = max[i];
            Float inv_d = 1.0f / d[i];
            Float t0 = (mn - o[i]) * inv_d;
            Float t1 = (mx - o[i]) * inv_d;
            if (t1 < t0) std::swap(t0, t1);
            if (t_range.start < t0) t_range.start = t0;
            if (t1 < t_range.end) t_range.end = t1;
            if (t_range.end <= t_range.start) return false;
        }
        return true;
    }
    int longest_axis() const {                                                              // aabb.rs:63-78
        Float s0 = max.x - min.x, s1 = max.y - min.y, s2 = max.z - min.z;
        if (s0 > s1) return (s0 > s2) ? 0 : 2;
        return (s1 > s2) ? 1 : 2;
    }
};
// f32::total_cmp (aabb.rs:80-82) as a strict-weak "less"
static inline bool total_less(Float a, Float b) {
    int32_t l, r; std::memcpy(&l, &a, 4); std::memcpy(&r, &b, 4);
    l ^= (int32_t)(((uint32_t)(l >> 31)) >> 1);
    r ^= (int32_t)(((uint32_t)(r >> 31)) >> 1);
    return l < r;
}

// ---------------------------------------------------------------- counters (SURVEY §8d event counts)
struct Counters {
    uint64_t rays = 0, node_visits = 0, sphere_tests = 0, quad_tests = 0, hits = 0, paths = 0;
    void add(const Counters& o) {
        rays += o.rays; node_visits += o.node_visits; sphere_tests += o.sphere_tests;
        quad_tests += o.quad_tests; hits += o.hits; paths += o.paths;
    }
};

// ---------------------------------------------------------------- material/mod.rs + hittable/mod.rs
struct Material;
struct HitRecord {                                                                          // hittable/mod.rs:19-25
    Float t; Vec3 point; Vec3 normal; bool front_face; const Material* material;
    uint32_t prim_id;   // not in the reference: insertion index, reported as the primary-hit AOV
};
static inline HitRecord make_hit(const Ray& ray, Float t, Vec3 outward_normal, const Material* m, uint32_t id) { // hittable/mod.rs:27-48
    HitRecord h;
    h.point = ray.at(t);
    h.front_face = dot(ray.direction, outward_normal) < 0.0f;
    h.normal = h.front_face ? normalized(outward_normal) : -normalized(outward_normal);
    h.t = t; h.material = m; h.prim_id = id;
    return h;
}
struct Material {                                                                           // material/mod.rs:5-11
    virtual ~Material() {}
    virtual bool scatter(const Ray& ray, const HitRecord& rec, Stream& rng, MathMode mode, Ray* out, Vec3* atten) const = 0;
    virtual bool emitted(Vec3* e) const { (void)e; return false; }
    virtual int kind() const = 0;
};
struct Lambertian : Material {                                                              // material/lambertian.rs:16-22
    Vec3 albedo;
    explicit Lambertian(Vec3 a) : albedo(a) {}
    bool scatter(const Ray&, const HitRecord& rec, Stream& rng, MathMode mode, Ray* out, Vec3* atten) const override {
        Vec3 dir = rec.normal + random_unit_vector(rng, mode);
        if (near_zero(dir)) dir = rec.normal;
        *out = make_ray(rec.point, dir); *atten = albedo; return true;
    }
    int kind() const override { return 0; }
};
struct Metal : Material {                                                                   // material/metal.rs:12-25
    Vec3 albedo; Float fuzz;
    Metal(Vec3 a, Float f) : albedo(a), fuzz(f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f)) {}     // clamp, metal.rs:13
    bool scatter(const Ray& ray, const HitRecord& rec, Stream& rng, MathMode mode, Ray* out, Vec3* atten) const override {
        Vec3 reflected = reflect(ray.direction, rec.normal);
        *out = make_ray(rec.point, reflected + fuzz * random_in_unit_sphere(rng, mode, true));
        *atten = albedo; return true;
    }
    int kind() const override { return 1; }
};
struct Dielectric : Material {                                                              // material/dielectric.rs:16-46
    Vec3 albedo; Float refraction_index;
    Dielectric(Vec3 a, Float ri) : albedo(a), refraction_index(ri) {}
    static Float reflectance(Float c, Float ri) {                                           // dielectric.rs:16-22
        Float one = 1.0f;
        Float sqrt_r0 = (one - ri) / (one + ri);
        Float r0 = sqrt_r0 * sqrt_r0;
        Float m = one - c;
        Float m2 = m * m;                // powi(5) == ((m*m)*(m*m))*m
        Float m5 = (m2 * m2) * m;
        return r0 + (one - r0) * m5;
    }
    bool scatter(const Ray& ray, const HitRecord& rec, Stream& rng, MathMode, Ray* out, Vec3* atten) const override {
        Float ri = rec.front_face ? 1.0f / refraction_index : refraction_index;
        Float c = fminf(-dot(rec.normal, ray.direction), 1.0f);
        Float s = std::sqrt(1.0f - c * c);
        bool total_reflection = ri * s > 1.0f;
        Float refl = reflectance(c, ri);
        // dielectric.rs:38 draws xi only when !total_reflection; keyed Philox makes the draw count irrelevant.
        Float xi = rng.random();
        Vec3 dir = (total_reflection || refl > xi) ? reflect(ray.direction, rec.normal)
                                                   : refract(ray.direction, rec.normal, ri);
        *out = make_ray(rec.point, dir); *atten = albedo; return true;
    }
    int kind() const override { return 2; }
};
struct Light : Material {                                                                   // material/light.rs:17-23
    Vec3 color;
    explicit Light(Vec3 c) : color(c) {}
    bool scatter(const Ray&, const HitRecord&, Stream&, MathMode, Ray*, Vec3*) const override { return false; }
    bool emitted(Vec3* e) const override { *e = color; return true; }
    int kind() const override { return 3; }
};

// ---------------------------------------------------------------- hittable/{sphere,quad}.rs
struct Hittable {
    virtual ~Hittable() {}
    virtual bool hit(const Ray& ray, Range t_range, HitRecord* rec, Counters* c) const = 0;
    virtual AABB bounding_box() const = 0;
};
struct Sphere : Hittable {
    Vec3 center; Float radius; AABB bbox; const Material* material; uint32_t id;
    Sphere(Vec3 c, Float r, const Material* m, uint32_t id_) : center(c), radius(r), material(m), id(id_) { // sphere.rs:16-25
        Vec3 rv = diag(r); bbox = AABB::make(c - rv, c + rv);
    }
    bool hit(const Ray& ray, Range t_range, HitRecord* rec, Counters* c) const override {   // sphere.rs:29-54
        if (c) c->sphere_tests++;
        Vec3 oc = ray.origin - center;
        Float a = squared_length(ray.direction);
        Float half_b = dot(oc, ray.direction);
        Float cc = squared_length(oc) - radius * radius;
        Float discriminant = half_b * half_b - a * cc;
        if (discriminant < 0.0f) return false;
        Float sqrtd = std::sqrt(discriminant);
        Float t = (-half_b - sqrtd) / a;
        if (!t_range.contains(t)) {
            t = (-half_b + sqrtd) / a;
            if (!t_range.contains(t)) return false;
        }
        Vec3 p = ray.at(t);
        *rec = make_hit(ray, t, p - center, material, id);
        return true;
    }
    AABB bounding_box() const override { return bbox; }
};
struct Quad : Hittable {
    Vec3 corner, u, v, n, w; Float d; AABB bbox; const Material* material; uint32_t id;
    Quad(Vec3 corner_, Vec3 u_, Vec3 v_, const Material* m, uint32_t id_) : corner(corner_), u(u_), v(v_), material(m), id(id_) { // quad.rs:20-29
        bbox = AABB::merge(AABB::make(corner, corner + u + v), AABB::make(corner + u, corner + v));
        n = cross(u, v);
        w = n / dot(n, n);
        d = dot(n, corner);
    }
    bool hit(const Ray& ray, Range t_range, HitRecord* rec, Counters* c) const override {   // quad.rs:33-54
        if (c) c->quad_tests++;
        Float dir_norm = dot(ray.direction, n);
        Float t = (d - dot(ray.origin, n)) / dir_norm;
        if (!t_range.contains(t)) return false;
        Vec3 p = ray.at(t) - corner;
        Float planar_x = dot(cross(p, v), w);
        Float planar_y = dot(cross(u, p), w);
        Range unit = {0.0f, 1.0f};
        if (!(unit.contains(planar_x) && unit.contains(planar_y))) return false;
        *rec = make_hit(ray, t, n, material, id);
        return true;
    }
    AABB bounding_box() const override { return bbox; }
};

// ---------------------------------------------------------------- hittable/bvh.rs
struct Node {
    std::unique_ptr<Node> left, right;
    const Hittable* hittable = nullptr;
    AABB bbox;
    uint32_t leaf_rank = 0;   // not in the reference: left-to-right leaf order, exported for tests

    static std::unique_ptr<Node> build(std::vector<const Hittable*>& objects, size_t lo, size_t hi, uint32_t* next_rank) { // bvh.rs:42-84
        size_t n = hi - lo;
        AABB bbox = objects[lo]->bounding_box();
        for (size_t i = lo + 1; i < hi; ++i) bbox = AABB::merge(bbox, objects[i]->bounding_box());
        int axis = bbox.longest_axis();
        std::unique_ptr<Node> node(new Node());
        if (n == 1) {
            node->hittable = objects[lo];
            node->bbox = objects[lo]->bounding_box();
            node->leaf_rank = (*next_rank)++;
        } else if (n == 2) {                                  // bvh.rs:58-68: no sort
            node->left = build(objects, lo, lo + 1, next_rank);
            node->right = build(objects, lo + 1, hi, next_rank);
            node->bbox = AABB::merge(node->left->bbox, node->right->bbox);
        } else {                                              // bvh.rs:69-83: stable sort on bbox.min[axis], total_cmp
            std::vector<const Hittable*> sub(objects.begin() + lo, objects.begin() + hi);   // `objects.to_vec()`
            std::stable_sort(sub.begin(), sub.end(), [axis](const Hittable* a, const Hittable* b) {
                return total_less(a->bounding_box().min[axis], b->bounding_box().min[axis]);
            });
            size_t mid = n / 2;
            node->left = build(sub, 0, mid, next_rank);
            node->right = build(sub, mid, n, next_rank);
            node->bbox = AABB::merge(node->left->bbox, node->right->bbox);
        }
        return node;
    }
    bool hit(const Ray& ray, Range t_range, HitRecord* rec, Counters* c) const {            // bvh.rs:88-107
        if (c) c->node_visits++;
        if (!bbox.intersect(ray, t_range)) return false;
        if (hittable) return hittable->hit(ray, t_range, rec, c);
        HitRecord lrec;
        if (left->hit(ray, t_range, &lrec, c)) {
            HitRecord rrec;
            Range narrowed = {t_range.start, lrec.t};
            if (right->hit(ray, narrowed, &rrec, c)) { *rec = rrec; return true; }
            *rec = lrec; return true;
        }
        return right->hit(ray, t_range, rec, c);
    }
};

// ---------------------------------------------------------------- hittable/world.rs
struct World {
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Hittable>> geometries;     // insertion order, world.rs:23-25
    std::unique_ptr<Node> root;                            // world.rs:43-45 get_bvh
    void build_bvh() {
        std::vector<const Hittable*> objs;
        for (auto& g : geometries) objs.push_back(g.get());
        uint32_t rank = 0;
        root = objs.empty() ? nullptr : Node::build(objs, 0, objs.size(), &rank);
    }
    bool hit(const Ray& ray, Range r, HitRecord* rec, Counters* c) const {                  // bvh.rs:25-27
        return root ? root->hit(ray, r, rec, c) : false;
    }
    // brute force over every primitive in insertion order with the same narrowing rule; used by tests only
    bool hit_bruteforce(const Ray& ray, Range r, HitRecord* rec) const {
        bool any = false;
        for (auto& g : geometries) {
            HitRecord h;
            if (g->bounding_box().intersect(ray, r) && g->hit(ray, r, &h, nullptr)) { any = true; *rec = h; r.end = h.t; }
        }
        return any;
    }
};

// ---------------------------------------------------------------- camera.rs
struct Camera {
    Vec3 position, viewport_upper_left, forward, horizontal, vertical, defocus_disk_u, defocus_disk_v;
    uint32_t width, height;
    static Camera make(Float focus_distance, Float defocus_angle, Vec3 position, Vec3 look_at, Vec3 up,
                       Float vertical_fov, uint32_t width, uint32_t height) {               // camera.rs:17-56
        const Float RADS_PER_DEG = 3.14159265358979323846f / 180.0f;                        // f32::to_radians
        Camera c;
        Float viewport_height = 2.0f * focus_distance * std::tan((vertical_fov * RADS_PER_DEG) / 2.0f);
        Float aspect_ratio = (Float)width / (Float)height;
        Float viewport_width = aspect_ratio * viewport_height;
        Vec3 w = normalized(position - look_at);
        Vec3 u = normalized(cross(up, w));
        Vec3 v = normalized(cross(w, u));
        c.forward = w * focus_distance;
        c.horizontal = u * viewport_width;
        c.vertical = v * viewport_height;
        c.viewport_upper_left = position - c.horizontal / 2.0f + c.vertical / 2.0f - c.forward;
        Float defocus_radius = focus_distance * std::tan((defocus_angle * RADS_PER_DEG) / 2.0f);
        c.defocus_disk_u = u * defocus_radius;
        c.defocus_disk_v = v * defocus_radius;
        c.position = position; c.width = width; c.height = height;
        return c;
    }
    Ray get_ray(Float u, Float v, Stream& rng) const {                                      // camera.rs:58-66
        Vec3 p = random_in_unit_disk(rng);
        Vec3 origin = position + p.x * defocus_disk_u + p.y * defocus_disk_v;
        return make_ray(origin, viewport_upper_left + (u * horizontal) - (v * vertical) - origin);
    }
};

// ---------------------------------------------------------------- renderer/{pointgen,sampler/cpu,imager}.rs
struct RenderParams {
    uint32_t width, height, spp, sample_offset, max_bounces;
    Float background[3];
    uint64_t seed;
    Float t_min;        // 0.001 in the reference (cpu.rs:48)
    uint32_t spp_total; // divisor of the accumulated sum (== spp for a whole render)
};

// pointgen.rs:37-52 for one (x, y, sample): jitter then Camera::get_ray
static inline Ray primary_ray(const Camera& cam, uint32_t x, uint32_t y, Stream& rng) {
    Float u = ((Float)x + rng.random()) / (Float)(cam.width - 1);
    Float v = ((Float)y + rng.random()) / (Float)(cam.height - 1);
    return cam.get_ray(u, v, rng);
}

// cpu.rs:39-65. `first` receives the primary hit (prim_id, t) when non-null.
static inline Vec3 single_point_sampling(const World& world, Ray ray, const RenderParams& p, uint32_t pixel, uint32_t sample,
                                         MathMode mode, Counters* c, uint32_t* first_prim, Float* first_t, uint32_t* n_rays) {
    uint32_t remain_bounces = p.max_bounces;
    Vec3 color = v3(0, 0, 0);
    Vec3 cumulated_attenuation = diag(1.0f);
    Vec3 background = v3(p.background[0], p.background[1], p.background[2]);
    uint32_t k = 0, rays = 0;
    if (first_prim) { *first_prim = 0xFFFFFFFFu; }
    if (first_t) { *first_t = std::numeric_limits<Float>::infinity(); }
    while (remain_bounces > 0) {
        HitRecord rec;
        Range r = {p.t_min, std::numeric_limits<Float>::infinity()};
        rays++;
        if (c) c->rays++;
        if (world.hit(ray, r, &rec, c)) {
            if (c) c->hits++;
            if (rays == 1) { if (first_prim) *first_prim = rec.prim_id; if (first_t) *first_t = rec.t; }
            Vec3 emission;
            if (!rec.material->emitted(&emission)) emission = v3(0, 0, 0);
            color = color + cumulated_attenuation * emission;
            k++;
            Stream rng(p.seed, pixel, sample, k);
            Ray new_ray; Vec3 attenuation;
            if (rec.material->scatter(ray, rec, rng, mode, &new_ray, &attenuation)) {
                cumulated_attenuation = cumulated_attenuation * attenuation;
                ray = new_ray;
                remain_bounces -= 1;
            } else break;
        } else {
            color = color + cumulated_attenuation * background;
            break;
        }
    }
    if (c) c->paths++;
    if (n_rays) *n_rays = rays;
    return color;
}

// utils/image.rs:92-111: gamma 2.2 then clamp [0,0.999] * 255 truncated (NaN -> 0 like Rust's saturating `as u8`)
static inline uint8_t tonemap_channel(Float lin) {
    Float g = std::pow(lin, 1.0f / 2.2f);
    Float cl = g < 0.0f ? 0.0f : (g > 0.999f ? 0.999f : g);   // f32::clamp; NaN stays NaN
    Float s = cl * 255.0f;
    if (!(s == s)) return 0;
    return (uint8_t)s;
}

} // namespace orc
