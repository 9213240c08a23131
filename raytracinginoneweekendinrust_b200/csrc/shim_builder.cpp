// shim_builder.cpp — CUDA-free half of the C ABI: the constructors of the reference's
// Hittable / Material / Texture implementors recorded as POD, plus host helpers that mirror
// reference functions (Tile::tile, Camera::new, hrpp::hash, write_ppm).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "shim_device.h"
#include "shim_internal.h"

using namespace shim;

static thread_local std::string g_err;
int shim::set_err(int code, const std::string& m) { g_err = m; return code; }

// ------------------------------------------------------------------------------------------
SHIM_API const char* shim_last_error(void) { return g_err.c_str(); }
SHIM_API int shim_version(void) { return 1; }

SHIM_API shim_scene* shim_scene_create(void) { return new (std::nothrow) shim_scene(); }
SHIM_API void shim_scene_destroy(shim_scene* s) {
    if (!s) return;
    if (s->dev) device_state_release(s->dev);
    delete s;
}

SHIM_API int shim_texture_solid(shim_scene* s, float r, float g, float b) {
    MUTABLE(s);
    HostTexture t; t.kind = TEX_SOLID; t.color[0] = r; t.color[1] = g; t.color[2] = b;
    s->sb.textures.push_back(t);
    return (int)s->sb.textures.size() - 1;
}
SHIM_API int shim_texture_checker(shim_scene* s, float scale, int even, int odd) {
    MUTABLE(s);
    if (!s->sb.ok_tex(even) || !s->sb.ok_tex(odd)) return set_err(SHIM_ERR_INVALID, "shim_texture_checker: bad texture id");
    HostTexture t; t.kind = TEX_CHECKER; t.scale = scale; t.even = even; t.odd = odd;
    s->sb.textures.push_back(t);
    return (int)s->sb.textures.size() - 1;
}
SHIM_API int shim_texture_marble(shim_scene* s, float scale, uint32_t seed) {
    MUTABLE(s);
    HostTexture t; t.kind = TEX_MARBLE; t.scale = scale; t.seed = seed;
    s->sb.textures.push_back(t);
    return (int)s->sb.textures.size() - 1;
}
SHIM_API int shim_texture_image(shim_scene* s, const uint8_t* rgb, int w, int h) {
    MUTABLE(s);
    if (!rgb || w <= 0 || h <= 0) return set_err(SHIM_ERR_INVALID, "shim_texture_image: bad image");
    HostTexture t; t.kind = TEX_IMAGE; t.w = w; t.h = h; t.rgb.assign(rgb, rgb + (size_t)w * h * 3);
    s->sb.textures.push_back(t);
    return (int)s->sb.textures.size() - 1;
}
static int push_material(shim_scene* s, const HostMaterial& m) { s->sb.materials.push_back(m); return (int)s->sb.materials.size() - 1; }
SHIM_API int shim_material_lambertian(shim_scene* s, int tex) {
    MUTABLE(s);
    if (!s->sb.ok_tex(tex)) return set_err(SHIM_ERR_INVALID, "shim_material_lambertian: bad texture id");
    HostMaterial m; m.kind = MAT_LAMBERTIAN; m.tex = tex; return push_material(s, m);
}
SHIM_API int shim_material_metal(shim_scene* s, float r, float g, float b, float fuzz) {
    MUTABLE(s);
    HostMaterial m; m.kind = MAT_METAL; m.albedo[0] = r; m.albedo[1] = g; m.albedo[2] = b;
    m.fuzz = fuzz < 0.0f ? 0.0f : (fuzz > 1.0f ? 1.0f : fuzz);  // metal.rs:17-22
    return push_material(s, m);
}
SHIM_API int shim_material_dielectric(shim_scene* s, float ior) {
    MUTABLE(s);
    HostMaterial m; m.kind = MAT_DIELECTRIC; m.ior = ior; return push_material(s, m);
}
SHIM_API int shim_material_diffuse_light(shim_scene* s, int tex) {
    MUTABLE(s);
    if (!s->sb.ok_tex(tex)) return set_err(SHIM_ERR_INVALID, "shim_material_diffuse_light: bad texture id");
    HostMaterial m; m.kind = MAT_DIFFUSE_LIGHT; m.tex = tex; return push_material(s, m);
}
SHIM_API int shim_material_isotropic(shim_scene* s, int tex) {
    MUTABLE(s);
    if (!s->sb.ok_tex(tex)) return set_err(SHIM_ERR_INVALID, "shim_material_isotropic: bad texture id");
    HostMaterial m; m.kind = MAT_ISOTROPIC; m.tex = tex; return push_material(s, m);
}

SHIM_API int shim_sphere(shim_scene* s, float cx, float cy, float cz, float r, int mat) {
    MUTABLE(s);
    if (!s->sb.ok_mat(mat)) return set_err(SHIM_ERR_INVALID, "shim_sphere: bad material id");
    HostHittable h; h.kind = H_SPHERE; h.p[0] = cx; h.p[1] = cy; h.p[2] = cz; h.p[3] = r; h.material = mat;
    return s->sb.add_hittable(h);
}
SHIM_API int shim_moving_sphere(shim_scene* s, float c0x, float c0y, float c0z, float c1x, float c1y, float c1z, float t0, float t1,
                                float r, int mat) {
    MUTABLE(s);
    if (!s->sb.ok_mat(mat)) return set_err(SHIM_ERR_INVALID, "shim_moving_sphere: bad material id");
    HostHittable h; h.kind = H_MSPHERE;
    const float v[9] = {c0x, c0y, c0z, c1x, c1y, c1z, t0, t1, r};
    memcpy(h.p, v, sizeof v);
    h.material = mat;
    return s->sb.add_hittable(h);
}
static int add_rect(shim_scene* s, int axis, float a0, float a1, float b0, float b1, float k, int mat) {
    if (!s->sb.ok_mat(mat)) return set_err(SHIM_ERR_INVALID, "rect: bad material id");
    HostHittable h; h.kind = H_RECT; h.axis = axis; h.p[0] = a0; h.p[1] = a1; h.p[2] = b0; h.p[3] = b1; h.p[4] = k; h.material = mat;
    return s->sb.add_hittable(h);
}
SHIM_API int shim_xy_rect(shim_scene* s, float x0, float x1, float y0, float y1, float z, int mat) { MUTABLE(s); return add_rect(s, 2, x0, x1, y0, y1, z, mat); }
SHIM_API int shim_xz_rect(shim_scene* s, float x0, float x1, float z0, float z1, float y, int mat) { MUTABLE(s); return add_rect(s, 1, x0, x1, z0, z1, y, mat); }
SHIM_API int shim_yz_rect(shim_scene* s, float y0, float y1, float z0, float z1, float x, int mat) { MUTABLE(s); return add_rect(s, 0, y0, y1, z0, z1, x, mat); }
SHIM_API int shim_tri(shim_scene* s, const float* p, int mat) {
    MUTABLE(s);
    if (!p || !s->sb.ok_mat(mat)) return set_err(SHIM_ERR_INVALID, "shim_tri: bad arguments");
    HostHittable h; h.kind = H_TRI; memcpy(h.p, p, 9 * sizeof(float)); h.material = mat;
    return s->sb.add_hittable(h);
}
SHIM_API int shim_cube(shim_scene* s, float x0, float y0, float z0, float x1, float y1, float z1, int mat) {
    MUTABLE(s);
    if (!s->sb.ok_mat(mat)) return set_err(SHIM_ERR_INVALID, "shim_cube: bad material id");
    HostHittable h; h.kind = H_CUBE; h.p[0] = x0; h.p[1] = y0; h.p[2] = z0; h.p[3] = x1; h.p[4] = y1; h.p[5] = z1; h.material = mat;
    return s->sb.add_hittable(h);
}
SHIM_API int shim_list_create(shim_scene* s) {
    MUTABLE(s);
    HostHittable h; h.kind = H_LIST; return s->sb.add_hittable(h);
}
SHIM_API int shim_list_add(shim_scene* s, int list, int hittable) {
    MUTABLE(s);
    if (!s->sb.ok_hit(list) || !s->sb.ok_hit(hittable) || s->sb.hittables[list].kind != H_LIST)
        return set_err(SHIM_ERR_INVALID, "shim_list_add: bad list or hittable id");
    s->sb.hittables[list].items.push_back(hittable);
    return SHIM_OK;
}
SHIM_API int shim_tris_bulk(shim_scene* s, const float* xyz, int n, int mat, int list) {
    MUTABLE(s);
    if (!xyz || n < 0 || !s->sb.ok_mat(mat) || !s->sb.ok_hit(list) || s->sb.hittables[list].kind != H_LIST)
        return set_err(SHIM_ERR_INVALID, "shim_tris_bulk: bad arguments");
    int first = (int)s->sb.hittables.size();
    s->sb.hittables.reserve(s->sb.hittables.size() + (size_t)n);
    for (int i = 0; i < n; ++i) {
        HostHittable h; h.kind = H_TRI; memcpy(h.p, xyz + (size_t)i * 9, 9 * sizeof(float)); h.material = mat;
        int id = s->sb.add_hittable(h);
        s->sb.hittables[list].items.push_back(id);
    }
    return first;
}
SHIM_API int shim_bvh(shim_scene* s, int list, float t0, float t1, uint64_t seed, int with_predictor) {
    MUTABLE(s);
    int id = s->sb.build_bvh(list, t0, t1, seed, with_predictor != 0);
    return id < 0 ? set_err(id, s->sb.err) : id;
}
SHIM_API int shim_bvh_from_nodes(shim_scene* s, int n, const int32_t* left, const int32_t* right, int root, float t0, float t1,
                                 int with_predictor) {
    MUTABLE(s);
    int id = s->sb.bvh_from_nodes(n, left, right, root, t0, t1, with_predictor != 0);
    return id < 0 ? set_err(id, s->sb.err) : id;
}
SHIM_API int shim_translate(shim_scene* s, int h, float dx, float dy, float dz) {
    MUTABLE(s);
    if (!s->sb.ok_hit(h)) return set_err(SHIM_ERR_INVALID, "shim_translate: bad hittable id");
    HostHittable t; t.kind = H_TRANSLATE; t.child = h; t.p[0] = dx; t.p[1] = dy; t.p[2] = dz;
    return s->sb.add_hittable(t);
}
SHIM_API int shim_rotate_y(shim_scene* s, int h, float degrees) {
    MUTABLE(s);
    if (!s->sb.ok_hit(h)) return set_err(SHIM_ERR_INVALID, "shim_rotate_y: bad hittable id");
    HostHittable r; r.kind = H_ROTATE_Y; r.child = h; r.p[0] = degrees;
    float radians = degrees * (3.14159265358979323846f / 180.0f);  // f32::to_radians, instance.rs:64
    r.sin_t = std::sin(radians); r.cos_t = std::cos(radians);
    return s->sb.add_hittable(r);
}
SHIM_API int shim_constant_medium(shim_scene* s, int boundary, float density, int tex) {
    MUTABLE(s);
    if (!s->sb.ok_hit(boundary) || !s->sb.ok_tex(tex)) return set_err(SHIM_ERR_INVALID, "shim_constant_medium: bad id");
    HostMaterial m; m.kind = MAT_ISOTROPIC; m.tex = tex;  // hittable.rs:156-158
    HostHittable c; c.kind = H_MEDIUM; c.child = boundary; c.phase_mat = push_material(s, m); c.neg_inv_density = -1.0f / density;
    return s->sb.add_hittable(c);
}
SHIM_API int shim_scene_set_option(shim_scene* s, int option, int value) {
    MUTABLE(s);
    if (option == SHIM_OPT_DEVICE_BVH) { s->sb.device_reference_topology = value == SHIM_DEVICE_BVH_REFERENCE; return SHIM_OK; }
    return set_err(SHIM_ERR_INVALID, "shim_scene_set_option: unknown option");
}
SHIM_API int shim_world_add(shim_scene* s, int h) {
    MUTABLE(s);
    if (!s->sb.ok_hit(h)) return set_err(SHIM_ERR_INVALID, "shim_world_add: bad hittable id");
    s->sb.world.push_back(h);
    return SHIM_OK;
}

SHIM_API int shim_bvh_info(shim_scene* s, int bvh, int* n_nodes, int* root, int* height) {
    NEED(s);
    if (!s->sb.ok_hit(bvh) || s->sb.hittables[bvh].kind != H_BVH) return set_err(SHIM_ERR_INVALID, "shim_bvh_info: not a bvh");
    const HostHittable& b = s->sb.hittables[bvh];
    if (n_nodes) *n_nodes = (int)b.nodes.size();
    if (root) *root = b.root;
    if (height) *height = b.height;
    return SHIM_OK;
}
SHIM_API int shim_bvh_nodes(shim_scene* s, int bvh, int32_t* left, int32_t* right, int32_t* parent, float* boxes) {
    NEED(s);
    if (!s->sb.ok_hit(bvh) || s->sb.hittables[bvh].kind != H_BVH) return set_err(SHIM_ERR_INVALID, "shim_bvh_nodes: not a bvh");
    const HostHittable& b = s->sb.hittables[bvh];
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        if (left) left[i] = b.nodes[i].left;
        if (right) right[i] = b.nodes[i].right;
        if (parent) parent[i] = b.nodes[i].parent;
        if (boxes) { memcpy(boxes + i * 6, b.nodes[i].box.mn, 12); memcpy(boxes + i * 6 + 3, b.nodes[i].box.mx, 12); }
    }
    return SHIM_OK;
}
SHIM_API uint64_t shim_scene_device_bytes(shim_scene* s) { return (s && s->committed) ? s->flat.bytes() : 0; }

// ------------------------------------------------------------------------------------------ host helpers
SHIM_API int shim_tile_layout(int W, int H, int tw, int th, int32_t* out, int cap) {
    if (W < 1 || H < 1 || tw < 1 || th < 1) return set_err(SHIM_ERR_INVALID, "shim_tile_layout: bad arguments");
    std::vector<TileRect> t = tile_layout(W, H, tw, th);
    for (size_t i = 0; i < t.size() && (int)i < cap && out; ++i) {
        out[i * 4 + 0] = t[i].width; out[i * 4 + 1] = t[i].height; out[i * 4 + 2] = t[i].x0; out[i * 4 + 3] = t[i].y0;
    }
    return (int)t.size();
}
SHIM_API int shim_camera_fields(const shim_camera* cam, float* out) {
    if (!cam || !out) return set_err(SHIM_ERR_INVALID, "shim_camera_fields: null argument");
    CameraPod c;
    camera_new(cam->look_from, cam->look_at, cam->view_up, cam->vertical_fov, cam->aspect_ratio, cam->aperture, cam->focus_dist,
               cam->time_start, cam->time_end, c);
    const f3* f[6] = {&c.origin, &c.horizontal, &c.vertical, &c.llc, &c.u, &c.v};
    for (int i = 0; i < 6; ++i) { out[i * 3] = f[i]->x; out[i * 3 + 1] = f[i]->y; out[i * 3 + 2] = f[i]->z; }
    out[18] = c.lens_radius; out[19] = c.time0; out[20] = c.time1;
    return SHIM_OK;
}
SHIM_API int shim_aabb_hit(const float* mn, const float* mx, const float* o, const float* d, float t_min, float t_max, int layout) {
    if (!mn || !mx || !o || !d) return set_err(SHIM_ERR_INVALID, "shim_aabb_hit: null argument");
    Ray r; r.o = mk3(o[0], o[1], o[2]); r.d = mk3(d[0], d[1], d[2]); r.time = 0;
    RayCtx c;
    make_ctx(c, r);
    float t_near;
    if (layout == 0) return slab(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], c, t_min, t_max, t_near) ? 1 : 0;
    // signed layout: the box as both children of an SNode, the ray picks the plane order by its direction signs
    f4 q[3];
    for (int k = 0; k < 3; ++k) {
        const bool neg = (c.soff[k] & 16) != 0;
        q[k] = neg ? f4{mx[k], mx[k], mn[k], mn[k]} : f4{mn[k], mn[k], mx[k], mx[k]};
    }
    float tl, tr; bool hl, hr;
    slab_pair_signed(q[0], q[1], q[2], c, t_min, t_max, tl, tr, hl, hr);
    return (hl && hr) ? 1 : ((hl || hr) ? set_err(SHIM_ERR_STATE, "shim_aabb_hit: the two children of one box disagree") : 0);
}
SHIM_API uint64_t shim_hrpp_hash(const float* o, const float* d) {
    Ray r; r.o = mk3(o[0], o[1], o[2]); r.d = mk3(d[0], d[1], d[2]); r.time = 0;
    return hrpp_hash(r);
}
SHIM_API int64_t shim_write_ppm(const float* rgb, int W, int H, const char* path) {
    if (!rgb || W < 1 || H < 1) return set_err(SHIM_ERR_INVALID, "shim_write_ppm: bad arguments");
    FILE* f = path ? fopen(path, "w") : stdout;
    if (!f) return set_err(SHIM_ERR_INVALID, "shim_write_ppm: cannot open file");
    int64_t bytes = fprintf(f, "P3\n%d %d\n255\n", W, H);
    for (int y = H - 1; y >= 0; --y)
        for (int x = 0; x < W; ++x) {
            int c[3];
            for (int k = 0; k < 3; ++k) {  // palette into_format::<u8>: clamp to [0,1], scale by 255, round to nearest
                float v = rgb[((size_t)y * W + x) * 3 + k];
                v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
                if (v != v) v = 0.0f;
                c[k] = (int)std::lround(v * 255.0f);
            }
            bytes += fprintf(f, "%d %d %d\n", c[0], c[1], c[2]);
        }
    if (path) fclose(f); else fflush(f);
    return bytes;
}
