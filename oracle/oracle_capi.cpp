// oracle/oracle_capi.cpp — C entry points over oracle.hpp so tests/ and bench.py can drive the CPU ORACLE through
// ctypes. TEST INFRASTRUCTURE ONLY (see the header of oracle.hpp). Multithreaded with std::thread over image rows
// (work-stealing by atomic row counter) — unlike the reference's single-producer/single-consumer channel pipeline
// (renderer/renderer.rs:62-77), so as a CPU baseline it is an upper bound on what the reference could reach.
#include "oracle.hpp"
#include <atomic>
#include <chrono>
#include <thread>
#include <mutex>

using namespace orc;

template <class F> static uint32_t parallel_rows(uint32_t height, uint32_t n_threads, F&& fn) {
    uint32_t nt = n_threads ? n_threads : std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    std::atomic<uint32_t> next(0);
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nt; ++t)
        th.emplace_back([&, t]() { for (;;) { uint32_t y = next.fetch_add(1); if (y >= height) break; fn(y, t); } });
    for (auto& x : th) x.join();
    return nt;
}


extern "C" {

struct orc_scene { World world; };

struct orc_camera {   // POD mirror of orc::Camera (7 Vec3 + w,h)
    float position[3], viewport_upper_left[3], forward[3], horizontal[3], vertical[3], defocus_disk_u[3], defocus_disk_v[3];
    uint32_t width, height;
};
struct orc_params {
    uint32_t width, height, spp, sample_offset, max_bounces;
    float background[3];
    uint64_t seed;
    float t_min;
    uint32_t spp_total;
    uint32_t math_mode;   // 0 = MATH_DET, 1 = MATH_LIBM
    uint32_t n_threads;   // 0 = hardware_concurrency
};
struct orc_stats { uint64_t paths, rays, node_visits, sphere_tests, quad_tests, hits; double wall_ms; uint32_t threads_used; };
struct orc_sample_point { uint32_t x, y; float origin[3], direction[3]; };   // renderer/pointgen.rs:7-13
struct orc_sampled_color { uint32_t x, y; float color[3]; };                 // renderer/imager.rs:9-15

static Vec3 f3(const float* p) { return v3(p[0], p[1], p[2]); }
static void s3(float* d, Vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

orc_scene* orc_scene_create() { return new orc_scene(); }
void orc_scene_destroy(orc_scene* s) { delete s; }

uint32_t orc_add_lambertian(orc_scene* s, const float* albedo) { s->world.materials.emplace_back(new Lambertian(f3(albedo))); return (uint32_t)s->world.materials.size() - 1; }
uint32_t orc_add_metal(orc_scene* s, const float* albedo, float fuzz) { s->world.materials.emplace_back(new Metal(f3(albedo), fuzz)); return (uint32_t)s->world.materials.size() - 1; }
uint32_t orc_add_dielectric(orc_scene* s, const float* albedo, float ior) { s->world.materials.emplace_back(new Dielectric(f3(albedo), ior)); return (uint32_t)s->world.materials.size() - 1; }
uint32_t orc_add_light(orc_scene* s, const float* color) { s->world.materials.emplace_back(new Light(f3(color))); return (uint32_t)s->world.materials.size() - 1; }

uint32_t orc_add_sphere(orc_scene* s, const float* c, float r, uint32_t mat) {
    uint32_t id = (uint32_t)s->world.geometries.size();
    s->world.geometries.emplace_back(new Sphere(f3(c), r, s->world.materials.at(mat).get(), id));
    return id;
}
uint32_t orc_add_quad(orc_scene* s, const float* corner, const float* u, const float* v, uint32_t mat) {
    uint32_t id = (uint32_t)s->world.geometries.size();
    s->world.geometries.emplace_back(new Quad(f3(corner), f3(u), f3(v), s->world.materials.at(mat).get(), id));
    return id;
}
// bulk: prim_type 0 = sphere (data: cx cy cz r), 1 = quad (data: corner u v = 9 floats); 12 floats per record
void orc_add_prims(orc_scene* s, uint32_t n, const uint32_t* type, const float* data12, const uint32_t* mat) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* d = data12 + 12 * (size_t)i;
        if (type[i] == 0) orc_add_sphere(s, d, d[3], mat[i]); else orc_add_quad(s, d, d + 3, d + 6, mat[i]);
    }
}
void orc_scene_commit(orc_scene* s) { s->world.build_bvh(); }
uint32_t orc_num_prims(orc_scene* s) { return (uint32_t)s->world.geometries.size(); }

static void collect_ranks(const Node* n, uint32_t* rank_of_prim) {
    if (!n) return;
    if (n->hittable) {
        const Sphere* sp = dynamic_cast<const Sphere*>(n->hittable);
        const Quad* q = dynamic_cast<const Quad*>(n->hittable);
        rank_of_prim[sp ? sp->id : q->id] = n->leaf_rank;
        return;
    }
    collect_ranks(n->left.get(), rank_of_prim); collect_ranks(n->right.get(), rank_of_prim);
}
// left-to-right leaf order of the reference-topology tree, indexed by prim_id
void orc_leaf_ranks(orc_scene* s, uint32_t* rank_of_prim) { collect_ranks(s->world.root.get(), rank_of_prim); }
static void tree_stats(const Node* n, uint32_t depth, uint32_t* nodes, uint32_t* maxd) {
    if (!n) return;
    (*nodes)++;
    if (depth > *maxd) *maxd = depth;
    tree_stats(n->left.get(), depth + 1, nodes, maxd); tree_stats(n->right.get(), depth + 1, nodes, maxd);
}
void orc_tree_stats(orc_scene* s, uint32_t* nodes, uint32_t* max_depth) { *nodes = 0; *max_depth = 0; tree_stats(s->world.root.get(), 1, nodes, max_depth); }

void orc_camera_make(float focus_distance, float defocus_angle, const float* position, const float* look_at, const float* up,
                     float vertical_fov, uint32_t width, uint32_t height, orc_camera* out) {
    Camera c = Camera::make(focus_distance, defocus_angle, f3(position), f3(look_at), f3(up), vertical_fov, width, height);
    s3(out->position, c.position); s3(out->viewport_upper_left, c.viewport_upper_left); s3(out->forward, c.forward);
    s3(out->horizontal, c.horizontal); s3(out->vertical, c.vertical); s3(out->defocus_disk_u, c.defocus_disk_u);
    s3(out->defocus_disk_v, c.defocus_disk_v); out->width = width; out->height = height;
}
static Camera cam_from(const orc_camera* c) {
    Camera r; r.position = f3(c->position); r.viewport_upper_left = f3(c->viewport_upper_left); r.forward = f3(c->forward);
    r.horizontal = f3(c->horizontal); r.vertical = f3(c->vertical); r.defocus_disk_u = f3(c->defocus_disk_u);
    r.defocus_disk_v = f3(c->defocus_disk_v); r.width = c->width; r.height = c->height; return r;
}
static RenderParams rp_from(const orc_params* p) {
    RenderParams r; r.width = p->width; r.height = p->height; r.spp = p->spp; r.sample_offset = p->sample_offset;
    r.max_bounces = p->max_bounces; r.background[0] = p->background[0]; r.background[1] = p->background[1];
    r.background[2] = p->background[2]; r.seed = p->seed; r.t_min = p->t_min; r.spp_total = p->spp_total ? p->spp_total : p->spp;
    return r;
}

// Whole render (pointgen.rs:37-52 -> cpu.rs:39-65 -> imager.rs:34-60 -> image.rs:92-111).
//   out_sum64   [H*W*3] optional: f64 sum of sample colours (samples sample_offset .. sample_offset+spp-1, in order)
//   out_linear  [H*W*3] optional: (float)(sum * (1.0/spp_total))
//   out_rgb8    [H*W*3] optional: tone-mapped
int orc_render(orc_scene* s, const orc_camera* cam_pod, const orc_params* p, double* out_sum64, float* out_linear,
               uint8_t* out_rgb8, orc_stats* stats) {
    Camera cam = cam_from(cam_pod); RenderParams rp = rp_from(p);
    MathMode mode = p->math_mode ? MATH_LIBM : MATH_DET;
    std::mutex mu; Counters total;
    auto t0 = std::chrono::steady_clock::now();
    uint32_t nt = parallel_rows(rp.height, p->n_threads, [&](uint32_t y, uint32_t) {
        Counters c;
        for (uint32_t x = 0; x < rp.width; ++x) {
            uint32_t pixel = y * rp.width + x;
            double acc[3] = {0, 0, 0};
            for (uint32_t si = 0; si < rp.spp; ++si) {
                uint32_t sample = rp.sample_offset + si;
                Stream rng(rp.seed, pixel, sample, 0);
                Ray ray = primary_ray(cam, x, y, rng);
                Vec3 col = single_point_sampling(s->world, ray, rp, pixel, sample, mode, &c, nullptr, nullptr, nullptr);
                acc[0] += (double)col.x; acc[1] += (double)col.y; acc[2] += (double)col.z;
            }
            double inv = 1.0 / (double)rp.spp_total;
            for (int ch = 0; ch < 3; ++ch) {
                size_t o = (size_t)pixel * 3 + ch;
                if (out_sum64) out_sum64[o] = acc[ch];
                float lin = (float)(acc[ch] * inv);
                if (out_linear) out_linear[o] = lin;
                if (out_rgb8) out_rgb8[o] = tonemap_channel(lin);
            }
        }
        std::lock_guard<std::mutex> g(mu); total.add(c);
    });
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->paths = total.paths; stats->rays = total.rays; stats->node_visits = total.node_visits;
        stats->sphere_tests = total.sphere_tests; stats->quad_tests = total.quad_tests; stats->hits = total.hits;
        stats->wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count(); stats->threads_used = nt;
    }
    return 0;
}

// Per-sample AOVs for samples sample_offset .. +spp-1 of every pixel, index (pixel*spp + si):
//   prim_id (0xFFFFFFFF = miss) and t of the primary hit, radiance[3] of the whole path, n_rays of the path.
int orc_render_aov(orc_scene* s, const orc_camera* cam_pod, const orc_params* p, uint32_t* prim_id, float* t,
                   float* radiance, uint32_t* n_rays) {
    Camera cam = cam_from(cam_pod); RenderParams rp = rp_from(p);
    MathMode mode = p->math_mode ? MATH_LIBM : MATH_DET;
    parallel_rows(rp.height, p->n_threads, [&](uint32_t y, uint32_t) {
        for (uint32_t x = 0; x < rp.width; ++x) {
            uint32_t pixel = y * rp.width + x;
            for (uint32_t si = 0; si < rp.spp; ++si) {
                uint32_t sample = rp.sample_offset + si;
                size_t o = (size_t)pixel * rp.spp + si;
                Stream rng(rp.seed, pixel, sample, 0);
                Ray ray = primary_ray(cam, x, y, rng);
                uint32_t fp, nr; Float ft;
                RenderParams q = rp;
                if (!radiance && !n_rays) q.max_bounces = q.max_bounces ? 1 : 0;   // primary hit only
                Vec3 col = single_point_sampling(s->world, ray, q, pixel, sample, mode, nullptr, &fp, &ft, &nr);
                if (prim_id) prim_id[o] = fp;
                if (t) t[o] = ft;
                if (radiance) { radiance[o * 3] = col.x; radiance[o * 3 + 1] = col.y; radiance[o * 3 + 2] = col.z; }
                if (n_rays) n_rays[o] = nr;
            }
        }
    });
    return 0;
}

// Trait-level adapter (renderer/sampler/mod.rs:9-16): SamplePoint stream in, SampledColor stream out, same order.
// RNG stream of item i: pixel = y*width + x, sample = sample_offset + i.
int orc_sample_batch(orc_scene* s, const orc_sample_point* in, size_t n, const orc_params* p, orc_sampled_color* out) {
    RenderParams rp = rp_from(p); MathMode mode = p->math_mode ? MATH_LIBM : MATH_DET;
    for (size_t i = 0; i < n; ++i) {
        Ray ray; ray.origin = f3(in[i].origin); ray.direction = f3(in[i].direction);
        uint32_t pixel = in[i].y * rp.width + in[i].x;
        Vec3 col = single_point_sampling(s->world, ray, rp, pixel, rp.sample_offset + (uint32_t)i, mode, nullptr, nullptr, nullptr, nullptr);
        out[i].x = in[i].x; out[i].y = in[i].y; s3(out[i].color, col);
    }
    return 0;
}

// Closest hit for arbitrary rays (directions used as given, NOT renormalised). use_bvh = 0 -> brute force.
int orc_trace_rays(orc_scene* s, size_t n, const float* origins, const float* dirs, float t_min, int use_bvh,
                   uint32_t* prim_id, float* t, float* normal) {
    for (size_t i = 0; i < n; ++i) {
        Ray ray; ray.origin = f3(origins + 3 * i); ray.direction = f3(dirs + 3 * i);
        HitRecord rec; Range r = {t_min, std::numeric_limits<Float>::infinity()};
        bool h = use_bvh ? s->world.hit(ray, r, &rec, nullptr) : s->world.hit_bruteforce(ray, r, &rec);
        prim_id[i] = h ? rec.prim_id : 0xFFFFFFFFu;
        t[i] = h ? rec.t : std::numeric_limits<Float>::infinity();
        if (normal) { Vec3 nn = h ? rec.normal : v3(0, 0, 0); s3(normal + 3 * i, nn); }
    }
    return 0;
}

// ---- unit-level entry points for KATs
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void orc_det_math(size_t n, const float* u, float* sin2pi, float* cos2pi, float* cbrt_out) {
    for (size_t i = 0; i < n; ++i) { det_sincos2pi(u[i], sin2pi + i, cos2pi + i); cbrt_out[i] = det_cbrt(u[i]); }
}
void orc_make_ray(const float* o, const float* d, float* out6) { Ray r = make_ray(f3(o), f3(d)); s3(out6, r.origin); s3(out6 + 3, r.direction); }
void orc_ray_at(const float* ray6, float t, float* out3) { Ray r; r.origin = f3(ray6); r.direction = f3(ray6 + 3); s3(out3, r.at(t)); }
void orc_vec3_kat(const float* a, const float* b, float* len_a, float* dot_ab, float* cross_ab) {
    *len_a = length(f3(a)); *dot_ab = dot(f3(a), f3(b)); s3(cross_ab, cross(f3(a), f3(b)));
}
// scatter one material directly: kind 0..3, params (albedo/colour[3], fuzz or ior); returns 1 if scattered
int orc_scatter(uint32_t kind, const float* albedo, float param, const float* ray6, float t, const float* outward_normal,
                uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t math_mode, float* out_ray6, float* out_atten) {
    std::unique_ptr<Material> m;
    if (kind == 0) m.reset(new Lambertian(f3(albedo))); else if (kind == 1) m.reset(new Metal(f3(albedo), param));
    else if (kind == 2) m.reset(new Dielectric(f3(albedo), param)); else m.reset(new Light(f3(albedo)));
    Ray r; r.origin = f3(ray6); r.direction = f3(ray6 + 3);
    HitRecord rec = make_hit(r, t, f3(outward_normal), m.get(), 0);
    Stream rng(seed, pixel, sample, bounce);
    Ray o; Vec3 a;
    if (!m->scatter(r, rec, rng, math_mode ? MATH_LIBM : MATH_DET, &o, &a)) return 0;
    s3(out_ray6, o.origin); s3(out_ray6 + 3, o.direction); s3(out_atten, a);
    return 1;
}
uint8_t orc_tonemap(float lin) { return tonemap_channel(lin); }
void orc_primary_ray(const orc_camera* cam_pod, uint32_t x, uint32_t y, uint64_t seed, uint32_t sample, float* out6) {
    Camera cam = cam_from(cam_pod); Stream rng(seed, y * cam.width + x, sample, 0);
    Ray r = primary_ray(cam, x, y, rng); s3(out6, r.origin); s3(out6 + 3, r.direction);
}

} // extern "C"
