// shimmer.hpp — header-only C++ mirror of the reference's construction API over the C ABI (shimmer_b200.h).
//
// Names and argument order follow the crate: Sphere::new(center, radius, material) (geometry/sphere.rs:26) becomes
// shimmer::Sphere::make(scene, center, radius, material) and so on; every object is a small handle (an id inside its
// Scene), so "Arc<dyn Hittable>" sharing is plain copying of handles.  Errors throw shimmer::Error carrying the
// status code and shim_last_error(); nothing is computed on the host.
#pragma once
#include <array>
#include <stdexcept>
#include <string>
#include <vector>

#include "shimmer_b200.h"

namespace shimmer {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
using Vec3 = std::array<float, 3>;

class Scene {
public:
    Scene() : s_(shim_scene_create()) { if (!s_) throw Error(SHIM_ERR_INVALID, "shim_scene_create failed"); }
    ~Scene() { shim_scene_destroy(s_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    shim_scene* raw() const { return s_; }
    int check(int rc) const { if (rc < 0) throw Error(rc, shim_last_error()); return rc; }
private:
    shim_scene* s_;
};

struct Texture { int id; };
struct Material { int id; };
struct Hittable { int id; };

// textures/*.rs
struct SolidColor { static Texture make(Scene& s, Vec3 c) { return {s.check(shim_texture_solid(s.raw(), c[0], c[1], c[2]))}; } };
struct Checker {
    static Texture make(Scene& s, float scale, Texture even, Texture odd) { return {s.check(shim_texture_checker(s.raw(), scale, even.id, odd.id))}; }
    static Texture from_color(Scene& s, float scale, Vec3 even, Vec3 odd) { return make(s, scale, SolidColor::make(s, even), SolidColor::make(s, odd)); }
};
struct Marble { static Texture make(Scene& s, float scale, uint32_t perlin_seed) { return {s.check(shim_texture_marble(s.raw(), scale, perlin_seed))}; } };
struct ImageTexture { static Texture make(Scene& s, const uint8_t* rgb8, int w, int h) { return {s.check(shim_texture_image(s.raw(), rgb8, w, h))}; } };

// materials/*.rs
struct Lambertian {
    static Material make(Scene& s, Texture albedo) { return {s.check(shim_material_lambertian(s.raw(), albedo.id))}; }
    static Material from_color(Scene& s, Vec3 albedo) { return make(s, SolidColor::make(s, albedo)); }
};
struct Metal { static Material make(Scene& s, Vec3 albedo, float fuzz) { return {s.check(shim_material_metal(s.raw(), albedo[0], albedo[1], albedo[2], fuzz))}; } };
struct Dialectric { static Material make(Scene& s, float index_of_refraction) { return {s.check(shim_material_dielectric(s.raw(), index_of_refraction))}; } };
struct DiffuseLight {
    static Material make(Scene& s, Texture emission) { return {s.check(shim_material_diffuse_light(s.raw(), emission.id))}; }
    static Material from_color(Scene& s, Vec3 c) { return make(s, SolidColor::make(s, c)); }
};
struct Isotropic {
    static Material make(Scene& s, Texture albedo) { return {s.check(shim_material_isotropic(s.raw(), albedo.id))}; }
    static Material from_color(Scene& s, Vec3 c) { return make(s, SolidColor::make(s, c)); }
};

// geometry/*.rs, hittable.rs, bvh.rs
struct Sphere { static Hittable make(Scene& s, Vec3 center, float radius, Material m) { return {s.check(shim_sphere(s.raw(), center[0], center[1], center[2], radius, m.id))}; } };
struct MovingSphere {
    static Hittable make(Scene& s, Vec3 c0, Vec3 c1, float t0, float t1, float radius, Material m) {
        return {s.check(shim_moving_sphere(s.raw(), c0[0], c0[1], c0[2], c1[0], c1[1], c1[2], t0, t1, radius, m.id))};
    }
};
struct XyRect { static Hittable make(Scene& s, float x0, float x1, float y0, float y1, float z, Material m) { return {s.check(shim_xy_rect(s.raw(), x0, x1, y0, y1, z, m.id))}; } };
struct XzRect { static Hittable make(Scene& s, float x0, float x1, float z0, float z1, float y, Material m) { return {s.check(shim_xz_rect(s.raw(), x0, x1, z0, z1, y, m.id))}; } };
struct YzRect { static Hittable make(Scene& s, float y0, float y1, float z0, float z1, float x, Material m) { return {s.check(shim_yz_rect(s.raw(), y0, y1, z0, z1, x, m.id))}; } };
struct Tri {
    static Hittable make(Scene& s, Vec3 p0, Vec3 p1, Vec3 p2, Material m) {
        float p[9] = {p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2]};
        return {s.check(shim_tri(s.raw(), p, m.id))};
    }
};
struct Cube { static Hittable make(Scene& s, Vec3 mn, Vec3 mx, Material m) { return {s.check(shim_cube(s.raw(), mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], m.id))}; } };
struct Translate { static Hittable make(Scene& s, Hittable h, Vec3 d) { return {s.check(shim_translate(s.raw(), h.id, d[0], d[1], d[2]))}; } };
struct RotateY { static Hittable make(Scene& s, Hittable h, float degrees) { return {s.check(shim_rotate_y(s.raw(), h.id, degrees))}; } };
struct ConstantMedium {
    static Hittable make(Scene& s, Hittable boundary, float density, Texture t) { return {s.check(shim_constant_medium(s.raw(), boundary.id, density, t.id))}; }
    static Hittable new_with_color(Scene& s, Hittable boundary, float density, Vec3 c) { return make(s, boundary, density, SolidColor::make(s, c)); }
};
class HittableList {
public:
    explicit HittableList(Scene& s) : s_(s), id_(s.check(shim_list_create(s.raw()))) {}
    void add(Hittable h) { s_.check(shim_list_add(s_.raw(), id_, h.id)); }
    Hittable as_hittable() const { return {id_}; }
private:
    Scene& s_;
    int id_;
};
struct Bvh {
    static Hittable make(Scene& s, const HittableList& list, float t0, float t1, uint64_t axis_seed = 0) { return {s.check(shim_bvh(s.raw(), list.as_hittable().id, t0, t1, axis_seed, 0))}; }
    static Hittable with_predictor(Scene& s, const HittableList& list, float t0, float t1, uint64_t axis_seed = 0) { return {s.check(shim_bvh(s.raw(), list.as_hittable().id, t0, t1, axis_seed, 1))}; }
};

// camera.rs:44-54
struct Camera {
    shim_camera pod;
    Camera(Vec3 from, Vec3 at, Vec3 vup, float vfov, float aspect, float aperture, float focus_dist, float t0, float t1) {
        pod = shim_camera{{from[0], from[1], from[2]}, {at[0], at[1], at[2]}, {vup[0], vup[1], vup[2]}, vfov, aspect, aperture, focus_dist, t0, t1};
    }
};

// renderer.rs:22-105
class Renderer {
public:
    Renderer(int image_width, int image_height) : w_(image_width), h_(image_height) {}
    static Renderer from_aspect_ratio(int image_width, float aspect_ratio) { return Renderer(image_width, (int)((float)image_width / aspect_ratio)); }
    int width() const { return w_; }
    int height() const { return h_; }
    /// `world` = the objects added with Scene/shim_world_add, in order.  Returns the linear-radiance image
    /// (row-major, y = 0 bottom); write it with shim_write_ppm for the reference's stdout format.
    std::vector<float> render(Scene& world, const Camera& camera, Vec3 background, uint32_t samples_per_pixel, uint32_t max_depth,
                              size_t tile_width = 8, size_t tile_height = 8, bool predictors = false, uint64_t seed = 0,
                              shim_stats* stats = nullptr) const {
        shim_render_params p{};
        p.width = w_; p.height = h_; p.samples_per_pixel = (int32_t)samples_per_pixel; p.max_depth = (int32_t)max_depth;
        p.tile_width = (int32_t)tile_width; p.tile_height = (int32_t)tile_height;
        p.background[0] = background[0]; p.background[1] = background[1]; p.background[2] = background[2];
        p.seed = seed; p.flags = predictors ? SHIM_RENDER_PREDICTORS : 0;
        std::vector<float> rgb((size_t)w_ * h_ * 3);
        shim_stats local;
        world.check(shim_render(world.raw(), &camera.pod, &p, rgb.data(), stats ? stats : &local));
        return rgb;
    }
private:
    int w_, h_;
};

}  // namespace shimmer
