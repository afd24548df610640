// examples/cornell_multi.cpp — BASELINE config 3 (Cornell box 1920x1080, 1024 spp, depth 50) on N GPUs of one box from a plain
// C++ host program: no Python, no NCCL, one call (trt::Renderer::render -> trt_render_multi). It renders the frame on one GPU
// too and reports how far the N-GPU image is from it (the contract: <= 1 ulp of f32, because the f64 partial sums are
// added in rank order instead of sample order) and the speed-up, i.e. the scaling efficiency THROUGH THE C ABI.
//   cornell_multi [n_gpus=all] [width=1920] [height=1080] [spp=1024] [out.png]
// Build:  g++ -std=c++17 -O2 -Iinclude examples/cornell_multi.cpp -o cornell_multi -Ltiny-raytracer_b200 -ltrt_b200 -Wl,-rpath,$PWD/tiny-raytracer_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "trt.hpp"

static trt::World build_world() {   // /root/reference/src/main.rs:23-87
    trt::World world;
    world.add_material("red", trt::Lambertian({0.65f, 0.05f, 0.05f}));
    world.add_material("white", trt::Lambertian({0.73f, 0.73f, 0.73f}));
    world.add_material("green", trt::Lambertian({0.12f, 0.45f, 0.15f}));
    world.add_material("light", trt::Light(trt::Vec3::new_diagonal(15.0f)));
    world.add_geometry(trt::Quad({100, 0, 0}, {0, 100, 0}, {0, 0, 100}, world.get_material("green")));
    world.add_geometry(trt::Quad({0, 0, 0}, {0, 100, 0}, {0, 0, 100}, world.get_material("red")));
    world.add_geometry(trt::Quad({65, 100, 60}, {-30, 0, 0}, {0, 0, -20}, world.get_material("light")));
    world.add_geometry(trt::Quad({0, 0, 0}, {100, 0, 0}, {0, 0, 100}, world.get_material("white")));
    world.add_geometry(trt::Quad({100, 100, 100}, {-100, 0, 0}, {0, 0, -100}, world.get_material("white")));
    world.add_geometry(trt::Quad({0, 0, 100}, {100, 0, 0}, {0, 100, 0}, world.get_material("white")));
    world.add_geometry(trt::new_box({25, 0, 50}, {55, 60, 80}, world.get_material("white")));
    world.add_geometry(trt::new_box({45, 0, 10}, {75, 30, 40}, world.get_material("white")));
    return world;
}

int main(int argc, char** argv) {
    try {
        int n = argc > 1 ? std::atoi(argv[1]) : 0;
        const uint32_t w = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 1920, h = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 1080;
        const uint32_t spp = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 1024;
        if (n <= 0) {                                   // all GPUs of the box: probe by creating contexts until one fails
            for (n = 0; n < 16; ++n) { trt_ctx* c = nullptr; if (trt_ctx_create(n, &c) != TRT_OK) break; trt_ctx_destroy(c); }
            if (n == 0) { std::fprintf(stderr, "trt error: %s\n", trt_last_error()); return 1; }
        }
        trt::World world = build_world();
        trt::Camera camera(140.0f, 0.6f, {50, 50, -140}, {50, 50, 0}, {0, 1, 0}, 40.0f, w, h);
        std::vector<int> devices; for (int i = 0; i < n; ++i) devices.push_back(i);
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };

        trt::Renderer one(spp, 8, 50, false, trt::Vec3::new_diagonal(0.001f));
        one.render(camera, world);                                         // warm-up (context creation, module load)
        auto t0 = now(); trt::Image img1 = one.render(camera, world); auto t1 = now();
        trt::Renderer many(spp, 8, 50, false, trt::Vec3::new_diagonal(0.001f));
        many.set_devices(devices);
        many.render(camera, world);                                        // warm-up: N contexts, peer mappings
        auto t2 = now(); trt::Image imgn = many.render(camera, world); auto t3 = now();

        long long max_ulp = 0, differing = 0; int max_lsb = 0;
        for (size_t i = 0; i < img1.linear.size(); ++i) {
            int32_t a, b; std::memcpy(&a, &img1.linear[i], 4); std::memcpy(&b, &imgn.linear[i], 4);
            long long d = a > b ? (long long)a - b : (long long)b - a;
            if (d > max_ulp) max_ulp = d;
            if (d) ++differing;
            int l = std::abs((int)img1.rgb8[i] - (int)imgn.rgb8[i]); if (l > max_lsb) max_lsb = l;
        }
        const trt_stats &s1 = one.last_stats, &sn = many.last_stats;
        std::printf("{\"example\": \"cornell_multi\", \"n_gpus\": %d, \"width\": %u, \"height\": %u, \"spp\": %u, \"rays_1gpu\": %llu, \"rays_ngpu\": %llu, "
                    "\"device_ms_1gpu\": %.3f, \"device_ms_ngpu\": %.3f, \"reduce_tonemap_ms_ngpu\": %.3f, \"wall_ms_1gpu\": %.2f, \"wall_ms_ngpu\": %.2f, "
                    "\"mrays_per_s_1gpu\": %.1f, \"mrays_per_s_ngpu\": %.1f, \"speedup_device\": %.3f, \"efficiency_device\": %.4f, "
                    "\"speedup_wall\": %.3f, \"linear_max_ulp_vs_1gpu\": %lld, \"linear_values_differing\": %lld, \"rgb8_max_lsb_vs_1gpu\": %d}\n",
                    n, w, h, spp, (unsigned long long)s1.rays, (unsigned long long)sn.rays, s1.device_ms, sn.device_ms, sn.finalize_ms, ms(t0, t1), ms(t2, t3),
                    s1.rays / (s1.device_ms * 1e3), sn.rays / ((sn.device_ms + sn.finalize_ms) * 1e3),
                    s1.device_ms / (sn.device_ms + sn.finalize_ms), s1.device_ms / (sn.device_ms + sn.finalize_ms) / n, ms(t0, t1) / ms(t2, t3),
                    max_ulp, differing, max_lsb);
        if (argc > 5) imgn.save(argv[5]);
        return (s1.rays == sn.rays && max_ulp <= 1 && max_lsb <= 1) ? 0 : 2;
    } catch (const trt::Error& e) {
        std::fprintf(stderr, "trt error %d: %s\n", e.code, e.what());
        return 1;
    }
}
