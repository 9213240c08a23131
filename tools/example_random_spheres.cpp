// Book-1 final scene through the C++ mirror (include/shimmer.hpp): the C++ twin of main.rs:185-251 + main.rs:105-183.
//   g++ -std=c++17 -Iinclude tools/example_random_spheres.cpp -Lraytracinginoneweekendinrust_b200/lib -lshimmer_b200 -o /tmp/ex
#include <cstdio>
#include <random>

#include "shimmer.hpp"

using namespace shimmer;

int main(int argc, char** argv) {
    try {
        Scene s;
        std::mt19937 rng(1);
        std::uniform_real_distribution<float> u(0.0f, 1.0f);
        HittableList world(s);
        world.add(Sphere::make(s, {0, -1000, 0}, 1000, Lambertian::make(s, Checker::from_color(s, 10.0f, {0.2f, 0.3f, 0.1f}, {0.9f, 0.9f, 0.9f}))));
        for (int a = -11; a < 11; ++a)
            for (int b = -11; b < 11; ++b) {
                float choose = u(rng);
                Vec3 c{a + 0.9f * u(rng), 0.2f, b + 0.9f * u(rng)};
                float dx = c[0] - 4.0f, dz = c[2];
                if (std::sqrt(dx * dx + dz * dz) <= 0.9f) continue;
                Material m = choose < 0.8f ? Lambertian::from_color(s, {u(rng) * u(rng), u(rng) * u(rng), u(rng) * u(rng)})
                           : choose < 0.95f ? Metal::make(s, {0.5f + 0.5f * u(rng), 0.5f + 0.5f * u(rng), 0.5f + 0.5f * u(rng)}, 0.5f * u(rng))
                                            : Dialectric::make(s, 1.5f);
                world.add(Sphere::make(s, c, 0.2f, m));
            }
        world.add(Sphere::make(s, {0, 1, 0}, 1.0f, Dialectric::make(s, 1.5f)));
        world.add(Sphere::make(s, {-4, 1, 0}, 1.0f, Lambertian::from_color(s, {0.4f, 0.2f, 0.1f})));
        world.add(Sphere::make(s, {4, 1, 0}, 1.0f, Metal::make(s, {0.7f, 0.6f, 0.5f}, 0.0f)));
        s.check(shim_world_add(s.raw(), Bvh::make(s, world, 0.0f, 1.0f, 1).id));
        s.check(shim_commit(s.raw()));
        Camera cam({13, 2, 3}, {0, 0, 0}, {0, 1, 0}, 20.0f, 1.5f, 0.1f, 10.0f, 0.0f, 0.0f);
        Renderer r = Renderer::from_aspect_ratio(argc > 1 ? atoi(argv[1]) : 600, 1.5f);
        shim_stats st;
        std::vector<float> rgb = r.render(s, cam, {0.7f, 0.8f, 1.0f}, 10, 50, 8, 8, false, 0, &st);
        shim_write_ppm(rgb.data(), r.width(), r.height(), argc > 2 ? argv[2] : nullptr);
        fprintf(stderr, "Render time (device): %.3f ms, %llu rays\n", st.device_ms, (unsigned long long)st.rays);
    } catch (const Error& e) {
        fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
