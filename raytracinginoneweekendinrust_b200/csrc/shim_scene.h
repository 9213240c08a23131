// shim_scene.h — host side of the boundary: records the reference's constructors as POD,
// builds the BVH the way bvh.rs does, and flattens the world into the SoA arrays the
// kernels read.  Pure C++ (no CUDA) so that it can be unit-tested without a device.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>
#include "shim_types.h"

namespace shim {

struct Box { float mn[3], mx[3]; };

enum HKind { H_SPHERE, H_MSPHERE, H_RECT, H_TRI, H_CUBE, H_LIST, H_BVH, H_TRANSLATE, H_ROTATE_Y, H_MEDIUM };

struct HostBvhNode { int left, right, parent; Box box; };  // child >= 0: node index; < 0: ~hittable id

struct HostTexture {
    int kind = TEX_SOLID;
    float color[3] = {0, 0, 0};
    float scale = 0;
    int even = -1, odd = -1;
    uint32_t seed = 0;
    std::vector<uint8_t> rgb;
    int w = 0, h = 0;
};
struct HostMaterial { int kind = MAT_LAMBERTIAN; int tex = -1; float albedo[3] = {0, 0, 0}; float fuzz = 0, ior = 1; };
struct HostHittable {
    int kind = H_SPHERE;
    float p[12] = {0};     // primitive parameters (see shim_scene.cpp)
    int axis = 0;          // rect
    int material = -1;
    int child = -1;        // transform / medium / bvh source list
    std::vector<int> items;          // list
    std::vector<HostBvhNode> nodes;  // bvh (post-order, root last when built here)
    int root = -1, height = 0;
    std::vector<HostBvhNode> dev_nodes;   // bvh: the binned-SAH tree the kernels walk, built with the BVH (Bvh::new's stand-in)
    int dev_root = -1, dev_height = 0;    // dev_root < 0: not built (unsupported member: reported at commit)
    float t0 = 0, t1 = 0;
    bool predictor = false;
    float sin_t = 0, cos_t = 1;      // rotate_y
    float neg_inv_density = 0;       // medium
    int phase_mat = -1;
};

// Flattened scene, host copy; uploaded verbatim.
struct FlatScene {
    std::vector<DevNode> nodes;
    std::vector<SNode> snodes;       // `nodes` in the signed layout (build_signed_nodes; empty = not built)
    void build_signed_nodes();
    std::vector<QNode> qnodes;       // the tree of object q_object quantised (build_quantized_nodes; empty = not built)
    int q_object = -1;
    bool build_quantized_nodes(int object_index);   // false: not representable (left empty)
    std::vector<double> sph;
    std::vector<f4> sph_s;
    std::vector<int> sph_mat;
    std::vector<f4> msph, rect, tri, cube;
    std::vector<DevObject> objects;
    std::vector<f4> materials, textures;
    std::vector<uint8_t> images, perlin;
    std::vector<int> handle[5], rank[5], leaf[5], sibling[5];
    std::vector<int> predictor_bvh;  // hittable id of each BVH that carries a predictor
    bool slow_lambertians = false;   // some material references carry the class MQ_SLOW_LAMBERTIAN (shim_types.h)
    SceneView view() const;          // pointers into the host vectors
    uint64_t bytes() const;
};

struct SceneBuilder {
    std::vector<HostTexture> textures;
    std::vector<HostMaterial> materials;
    std::vector<HostHittable> hittables;
    std::vector<int> world;
    std::string err;
    bool device_reference_topology = false;  // walk the recorded bvh.rs tree instead of the SAH rebuild

    bool ok_tex(int t) const { return t >= 0 && t < (int)textures.size(); }
    bool ok_mat(int m) const { return m >= 0 && m < (int)materials.size(); }
    bool ok_hit(int h) const { return h >= 0 && h < (int)hittables.size(); }

    int add_hittable(const HostHittable& h) { hittables.push_back(h); return (int)hittables.size() - 1; }
    // Hittable::bounding_box for the kinds a BVH may contain; false = unsupported
    bool bounding_box(int h, float t0, float t1, Box& out) const;
    // Bvh::new (bvh.rs:46-62, 249-333)
    int build_bvh(int list, float t0, float t1, uint64_t axis_seed, bool predictor);
    int bvh_from_nodes(int n, const int32_t* left, const int32_t* right, int root, float t0, float t1, bool predictor);
    // flatten the world; returns SHIM_OK or a negative status with `err` set
    int flatten(FlatScene& out);
};

// Camera::new, camera.rs:44-81
void camera_new(const float from[3], const float at[3], const float vup[3], float vfov, float aspect, float aperture,
                float focus_dist, float t0, float t1, CameraPod& out);
// Tile::tile, renderer.rs:242-296
struct TileRect { int width, height, x0, y0; };
std::vector<TileRect> tile_layout(int W, int H, int tw, int th);
// pixel indices (y*W + x) in the order the reference's tile loop visits them (renderer.rs:63-84),
// optionally only tiles with index % world == rank
std::vector<uint32_t> tile_pixel_order(int W, int H, int tw, int th, int rank, int world);
// Perlin permutation tables of one marble texture (19 x 256 bytes)
void marble_tables(uint32_t seed, std::vector<uint8_t>& out);

}  // namespace shim
