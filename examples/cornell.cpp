// examples/cornell.cpp — /root/reference/src/main.rs (the Cornell box script) written against include/trt.hpp.
// Build:  g++ -std=c++17 -O2 -Iinclude examples/cornell.cpp -o examples/cornell -Ltiny-raytracer_b200 -ltrt_b200 -Wl,-rpath,'$ORIGIN/../tiny-raytracer_b200'
#include <cstdio>
#include "trt.hpp"

static trt::World build_world() {   // main.rs:23-87
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
        trt::World world = build_world();
        trt::Camera camera(140.0f, 0.6f, {50, 50, -140}, {50, 50, 0}, {0, 1, 0}, 40.0f, 300, 300);   // main.rs:8-16
        trt::Renderer instance(300, 8, 20, true, trt::Vec3::new_diagonal(0.001f));                   // main.rs:17
        trt::Image image = instance.render(camera, world);
        image.save(argc > 1 ? argv[1] : "output.png");                                                  // main.rs:19-20
        std::printf("%llu paths, %llu rays, %.2f ms on the device (%.1f Mrays/s)\n", (unsigned long long)instance.last_stats.paths,
                    (unsigned long long)instance.last_stats.rays, instance.last_stats.device_ms,
                    instance.last_stats.rays / (instance.last_stats.device_ms * 1e3));
        return 0;
    } catch (const trt::Error& e) {
        std::fprintf(stderr, "trt error %d: %s\n", e.code, e.what());
        return 1;
    }
}
