// include/trt.hpp — header-only C++17 mirror of tiny-raytracer's public Rust API over the C ABI of trt.h.
//
// The reference is compiled Rust and this image has no Rust toolchain, so the host side above the C ABI is written
// here in C++ with the reference's names, argument meanings and error behaviour (the Rust crate a maintainer would
// add is kept as source under rust/). A reference scene script such as /root/reference/src/main.rs
// translates line by line:
//
//     trt::World world;                                                       // World::new()            hittable/world.rs:16-21
//     world.add_material("white", trt::Lambertian({0.73f, 0.73f, 0.73f}));    // add_material            hittable/world.rs:27-33
//     world.add_geometry(trt::Quad({0,0,0}, {100,0,0}, {0,0,100}, world.get_material("white")));   // hittable/world.rs:23-25
//     trt::Camera camera(140.f, 0.6f, {50,50,-140}, {50,50,0}, {0,1,0}, 40.f, 300, 300);            // camera.rs:17-26
//     trt::Renderer renderer(300, 8, 20, true, trt::Vec3::new_diagonal(0.001f));                    // renderer/renderer.rs:21-35
//     trt::Image image = renderer.render(camera, world);                                            // renderer/renderer.rs:37
//     image.save_ppm("output.ppm");
//
// Errors: the reference panics (`.expect`, world.rs:29-31); this mirror throws trt::Error carrying the C ABI's status and
// message, so nothing unwinds across the C boundary.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>
#include "trt.h"

namespace trt {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != TRT_OK) throw Error(rc, trt_last_error()); }

struct Vec3 {                                                   // math/vec3.rs:9-30
    float x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static Vec3 new_diagonal(float v) { return Vec3(v, v, v); }
    static Vec3 zero() { return Vec3(); }
    const float* data() const { return &x; }
};

struct Material {                                               // material/mod.rs:5-11
    enum Kind { LAMBERTIAN = 0, METAL = 1, DIELECTRIC = 2, LIGHT = 3 } kind;
    Vec3 rgb; float param;
};
inline std::shared_ptr<Material> Lambertian(Vec3 albedo) { return std::make_shared<Material>(Material{Material::LAMBERTIAN, albedo, 0.f}); }
inline std::shared_ptr<Material> Metal(Vec3 albedo, float fuzz) { return std::make_shared<Material>(Material{Material::METAL, albedo, fuzz}); }
inline std::shared_ptr<Material> Dielectric(Vec3 albedo, float refraction_index) { return std::make_shared<Material>(Material{Material::DIELECTRIC, albedo, refraction_index}); }
inline std::shared_ptr<Material> Light(Vec3 color) { return std::make_shared<Material>(Material{Material::LIGHT, color, 0.f}); }

struct Geometry {                                               // hittable/{sphere,quad}.rs
    bool is_quad; Vec3 a, u, v; float radius; std::shared_ptr<Material> material;
};
inline Geometry Sphere(Vec3 center, float radius, std::shared_ptr<Material> m) { return Geometry{false, center, Vec3(), Vec3(), radius, std::move(m)}; }
inline Geometry Quad(Vec3 corner, Vec3 u, Vec3 v, std::shared_ptr<Material> m) { return Geometry{true, corner, u, v, 0.f, std::move(m)}; }

// new_box of /root/reference/src/main.rs:89-125 (front, right, back, left, top, bottom), flattened to six quads
inline std::vector<Geometry> new_box(Vec3 a, Vec3 b, std::shared_ptr<Material> m) {
    auto mn = [](float p, float q) { return p < q ? p : q; };
    auto mx = [](float p, float q) { return p > q ? p : q; };
    Vec3 lo(mn(a.x, b.x), mn(a.y, b.y), mn(a.z, b.z)), hi(mx(a.x, b.x), mx(a.y, b.y), mx(a.z, b.z));
    Vec3 dx(hi.x - lo.x, 0, 0), dy(0, hi.y - lo.y, 0), dz(0, 0, hi.z - lo.z), ndx(-dx.x, 0, 0), ndz(0, 0, -dz.z);
    return {Quad({lo.x, lo.y, hi.z}, dx, dy, m), Quad({hi.x, lo.y, hi.z}, ndz, dy, m), Quad({hi.x, lo.y, lo.z}, ndx, dy, m),
            Quad({lo.x, lo.y, lo.z}, dz, dy, m), Quad({lo.x, hi.y, hi.z}, dx, ndz, m), Quad({lo.x, lo.y, lo.z}, dx, dz, m)};
}

// math::transform::Transform (math/transform.rs:10-111): 4x4 f32 matrix, new(translation, rotation, scale) = T * S * (Rz * Ry * Rx).
// Dead code in the reference; here it also places instances: World::add_instance(geometries, transform) adds transformed
// copies (quad: corner through the matrix, edges through its linear part; sphere: similarity transforms only).
struct Transform {
    float m[4][4];
    Transform() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = i == j ? 1.f : 0.f; }
    static Transform identity() { return Transform(); }
    static Transform translate(Vec3 d) { Transform t; t.m[0][3] = d.x; t.m[1][3] = d.y; t.m[2][3] = d.z; return t; }
    static Transform scale(Vec3 s) { Transform t; t.m[0][0] = s.x; t.m[1][1] = s.y; t.m[2][2] = s.z; return t; }
    static Transform rotate_x(float a) { Transform t; float c = std::cos(a), s = std::sin(a); t.m[1][1] = c; t.m[1][2] = -s; t.m[2][1] = s; t.m[2][2] = c; return t; }
    static Transform rotate_y(float a) { Transform t; float c = std::cos(a), s = std::sin(a); t.m[0][0] = c; t.m[0][2] = s; t.m[2][0] = -s; t.m[2][2] = c; return t; }
    static Transform rotate_z(float a) { Transform t; float c = std::cos(a), s = std::sin(a); t.m[0][0] = c; t.m[0][1] = -s; t.m[1][0] = s; t.m[1][1] = c; return t; }
    static Transform rotate(Vec3 a) { return rotate_z(a.z) * rotate_y(a.y) * rotate_x(a.x); }
    static Transform make(Vec3 translation, Vec3 rotation, Vec3 scales) { return translate(translation) * scale(scales) * rotate(rotation); }   // Transform::new
    friend Transform operator*(const Transform& a, const Transform& b) {
        Transform r;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
        return r;
    }
    Vec3 operator*(Vec3 v) const { return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z + m[0][3], m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z + m[1][3],
                                               m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z + m[2][3]); }
    Vec3 linear(Vec3 v) const { return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z, m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z); }
    Geometry apply(const Geometry& g) const {
        if (g.is_quad) return Geometry{true, (*this) * g.a, linear(g.u), linear(g.v), 0.f, g.material};
        double k2 = 0; bool sim = true;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double d = 0; for (int k = 0; k < 3; ++k) d += (double)m[k][i] * m[k][j];
            if (i == 0 && j == 0) k2 = d;
            if (std::fabs(d - (i == j ? k2 : 0.0)) > 1e-5 * (k2 > 0 ? k2 : 1.0)) sim = false;
        }
        if (!sim || k2 <= 0) throw Error(TRT_ERR_INVALID, "Transform::apply: a sphere needs a similarity transform");
        return Geometry{false, (*this) * g.a, Vec3(), Vec3(), g.radius * (float)std::sqrt(k2), g.material};
    }
};

class World {                                                   // hittable/world.rs:10-55
public:
    void add_geometry(Geometry g) { geometries_.push_back(std::move(g)); }
    void add_geometry(const std::vector<Geometry>& list) { for (auto& g : list) geometries_.push_back(g); }
    // instanced geometry: transformed copies of a template (math/transform.rs)
    void add_instance(const std::vector<Geometry>& list, const Transform& t) { for (auto& g : list) geometries_.push_back(t.apply(g)); }
    void add_instance(const Geometry& g, const Transform& t) { geometries_.push_back(t.apply(g)); }
    void add_material(const std::string& name, std::shared_ptr<Material> m) {
        if (materials_.count(name)) throw Error(TRT_ERR_INVALID, name + " key is already in the material table");   // world.rs:29-31
        materials_[name] = std::move(m);
    }
    std::shared_ptr<Material> get_material(const std::string& name) const {
        auto it = materials_.find(name); return it == materials_.end() ? nullptr : it->second;
    }
    const std::vector<Geometry>& geometries() const { return geometries_; }
private:
    std::vector<Geometry> geometries_;
    std::map<std::string, std::shared_ptr<Material>> materials_;
};

class Camera {                                                  // camera.rs:17-56
public:
    Camera(float focus_distance, float defocus_angle, Vec3 position, Vec3 look_at, Vec3 up, float vertical_fov, uint32_t width, uint32_t height) {
        check(trt_camera_make(focus_distance, defocus_angle, position.data(), look_at.data(), up.data(), vertical_fov, width, height, &pod));
    }
    std::pair<uint32_t, uint32_t> get_image_size() const { return {pod.width, pod.height}; }
    trt_camera pod;
};

struct Image {                                                  // utils/image.rs
    uint32_t width = 0, height = 0;
    std::vector<float> linear;        // H*W*3, sum / spp
    std::vector<uint8_t> rgb8;        // gamma 2.2, clamp, * 255
    void save_ppm(const std::string& filename) const {
        FILE* f = std::fopen(filename.c_str(), "wb");
        if (!f) throw Error(TRT_ERR_INVALID, "cannot open " + filename);
        std::fprintf(f, "P6\n%u %u\n255\n", width, height);
        std::fwrite(rgb8.data(), 1, rgb8.size(), f); std::fclose(f);
    }
    // Image::save, utils/image.rs:66-90 (the reference writes an 8-bit RGB PNG through the `image` crate). Self-contained:
    // one IDAT of stored (uncompressed) deflate blocks, so no zlib is needed; any PNG reader decodes it.
    void save(const std::string& filename) const {
        std::vector<uint8_t> raw;                               // filter byte 0 + one RGB scanline, per row
        raw.reserve((size_t)height * (3 * (size_t)width + 1));
        for (uint32_t y = 0; y < height; ++y) {
            raw.push_back(0);
            raw.insert(raw.end(), rgb8.begin() + 3 * (size_t)width * y, rgb8.begin() + 3 * (size_t)width * (y + 1));
        }
        std::vector<uint8_t> z = {0x78, 0x01};
        uint32_t a = 1, b = 0;                                  // adler32 of raw
        for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
        for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
            size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
            z.push_back(off + n >= raw.size() ? 1 : 0);
            z.push_back(n & 255); z.push_back(n >> 8); z.push_back(~n & 255); z.push_back((~n >> 8) & 255);
            z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
            if (raw.empty()) break;
        }
        uint32_t adler = (b << 16) | a;
        for (int i = 3; i >= 0; --i) z.push_back((adler >> (8 * i)) & 255);
        FILE* f = std::fopen(filename.c_str(), "wb");
        if (!f) throw Error(TRT_ERR_INVALID, "cannot open " + filename);
        auto be32 = [](uint32_t v, uint8_t* o) { o[0] = v >> 24; o[1] = v >> 16; o[2] = v >> 8; o[3] = v; };
        auto chunk = [&](const char tag[5], const std::vector<uint8_t>& data) {
            std::vector<uint8_t> c(8 + data.size() + 4);
            be32((uint32_t)data.size(), c.data());
            for (int i = 0; i < 4; ++i) c[4 + i] = (uint8_t)tag[i];
            std::copy(data.begin(), data.end(), c.begin() + 8);
            uint32_t crc = 0xFFFFFFFFu;                          // crc32 over tag + data
            for (size_t i = 4; i < 8 + data.size(); ++i) {
                crc ^= c[i];
                for (int k = 0; k < 8; ++k) crc = (crc >> 1) ^ (0xEDB88320u & (0u - (crc & 1u)));
            }
            be32(~crc, c.data() + 8 + data.size());
            std::fwrite(c.data(), 1, c.size(), f);
        };
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
        std::fwrite(sig, 1, 8, f);
        std::vector<uint8_t> ihdr(13);
        be32(width, ihdr.data()); be32(height, ihdr.data() + 4);
        ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
        chunk("IHDR", ihdr); chunk("IDAT", z); chunk("IEND", {});
        std::fclose(f);
    }
};

class Renderer {                                                // renderer/renderer.rs:13-79
public:
    Renderer(uint32_t samples_per_pixel, uint32_t /*num_sampler_threads*/, uint32_t max_bounces, bool /*progressbar*/,
             std::optional<Vec3> background_color = std::nullopt, uint64_t seed = 1, int device = 0,
             int bvh_kind = TRT_BVH_AUTO, uint32_t variant = TRT_VARIANT_MEGAKERNEL)
        : spp_(samples_per_pixel), max_bounces_(max_bounces), bg_(background_color.value_or(Vec3::zero())), seed_(seed),
          device_(device), bvh_kind_(bvh_kind), variant_(variant) {}

    // Render on several GPUs of the box (trt_render_multi: samples-per-pixel partitioned, NVLink peer reduce + tone map on
    // the GPUs). Empty or one entry = the single-GPU path on that device.
    Renderer& set_devices(std::vector<int> devices) { devices_ = std::move(devices); return *this; }

    Image render(const Camera& camera, const World& world) {
        trt_ctx* ctx = nullptr; trt_scene* scene = nullptr;
        const bool multi = devices_.size() > 1;
        if (!multi) check(trt_ctx_create(devices_.empty() ? device_ : devices_[0], &ctx));
        try {
            check(trt_scene_create(&scene));
            std::map<const Material*, uint32_t> ids;
            auto mat_id = [&](const std::shared_ptr<Material>& m) {
                auto it = ids.find(m.get()); if (it != ids.end()) return it->second;
                uint32_t id = 0;
                switch (m->kind) {
                    case Material::LAMBERTIAN: check(trt_add_lambertian(scene, m->rgb.data(), &id)); break;
                    case Material::METAL: check(trt_add_metal(scene, m->rgb.data(), m->param, &id)); break;
                    case Material::DIELECTRIC: check(trt_add_dielectric(scene, m->rgb.data(), m->param, &id)); break;
                    default: check(trt_add_light(scene, m->rgb.data(), &id));
                }
                return ids[m.get()] = id;
            };
            for (const Geometry& g : world.geometries()) {
                if (!g.material) throw Error(TRT_ERR_INVALID, "geometry without material");
                uint32_t m = mat_id(g.material);
                if (g.is_quad) check(trt_add_quad(scene, g.a.data(), g.u.data(), g.v.data(), m, nullptr));
                else check(trt_add_sphere(scene, g.a.data(), g.radius, m, nullptr));
            }
            check(trt_scene_commit(scene, bvh_kind_));
            trt_params p{};
            p.width = camera.pod.width; p.height = camera.pod.height; p.spp = spp_; p.sample_offset = 0; p.max_bounces = max_bounces_;
            p.background[0] = bg_.x; p.background[1] = bg_.y; p.background[2] = bg_.z;
            p.seed = seed_; p.t_min = 0.001f; p.spp_total = spp_; p.variant = variant_; p.flags = 0;
            Image img; img.width = p.width; img.height = p.height;
            img.linear.resize((size_t)p.width * p.height * 3); img.rgb8.resize(img.linear.size());
            if (multi) check(trt_render_multi(devices_.data(), (int)devices_.size(), scene, &camera.pod, &p, img.linear.data(), img.rgb8.data(), &last_stats));
            else check(trt_render(ctx, scene, &camera.pod, &p, img.linear.data(), img.rgb8.data(), &last_stats));
            trt_scene_destroy(scene); trt_ctx_destroy(ctx);
            return img;
        } catch (...) { trt_scene_destroy(scene); trt_ctx_destroy(ctx); throw; }
    }
    trt_stats last_stats{};
private:
    uint32_t spp_, max_bounces_; Vec3 bg_; uint64_t seed_; int device_; int bvh_kind_; uint32_t variant_;
    std::vector<int> devices_;
};

} // namespace trt
