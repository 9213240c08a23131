// hostsim.cpp — TEST HARNESS ONLY.  Compiles the product's __host__ __device__ arithmetic
// (csrc/shim_device.h) and its flattener for the CPU, so that the device math can be checked
// against the oracle in this GPU-less container before spending GPU time.  It is not part
// of the package, is never loaded by it, and exists only under tests/.
//
// It links the product's CUDA-free builder entry points (shim_builder.cpp, shim_scene.cpp),
// so scenes are recorded through the very same shim_* calls; only commit / trace / radiance
// are replaced by serial loops over the shared inline functions.
#include <cstring>
#include <vector>

#include "../../raytracinginoneweekendinrust_b200/csrc/shim_device.h"
#include "../../raytracinginoneweekendinrust_b200/csrc/shim_internal.h"

using namespace shim;

void shim::device_state_release(DeviceState*) {}

// HRPP tables of the harness (the product keeps them in device memory)
#include <map>
namespace {
struct HrppHost { std::vector<HrppSlot> slots; int log2 = 0; bool on = false; uint64_t tp = 0, fp = 0, none = 0; };
std::map<shim_scene*, HrppHost> g_hrpp;
SceneView view_of(shim_scene* s) {
    SceneView sv = s->flat.view();
    auto it = g_hrpp.find(s);
    if (it != g_hrpp.end() && it->second.on) {
        sv.hrpp_slots = it->second.slots.data();
        sv.hrpp_log2 = it->second.log2; sv.hrpp_mask = (uint32_t)(((size_t)1 << it->second.log2) - 1);
    }
    return sv;
}
bool hrpp_on(shim_scene* s) { auto it = g_hrpp.find(s); return it != g_hrpp.end() && it->second.on; }
void hrpp_count(shim_scene* s, const TraceCounters& tc) {
    auto it = g_hrpp.find(s);
    if (it != g_hrpp.end()) { it->second.tp += tc.hrpp_tp; it->second.fp += tc.hrpp_fp; it->second.none += tc.hrpp_none; }
}
}  // namespace

// fresh predictor tables for every BVH recorded with a predictor (2^log2 slots each)
extern "C" __attribute__((visibility("default"))) int hs_enable_predictors(shim_scene* s, int log2) {
    HrppHost& h = g_hrpp[s];
    size_t n = s->flat.predictor_bvh.size();
    h.log2 = log2; h.on = n > 0; h.tp = h.fp = h.none = 0;
    HrppSlot empty;
    memset(&empty, 0xFF, sizeof empty);
    h.slots.assign(((size_t)1 << log2) * n, empty);
    return (int)n;
}
extern "C" __attribute__((visibility("default"))) void hs_predictor_stats(shim_scene* s, uint64_t* out3) {
    HrppHost& h = g_hrpp[s];
    out3[0] = h.tp; out3[1] = h.fp; out3[2] = h.none;
}

extern "C" __attribute__((visibility("default"))) int hs_commit(shim_scene* s);
// the harness library answers shim_commit with the CPU-side flatten only (no device)
extern "C" __attribute__((visibility("default"))) int shim_commit(shim_scene* s) { return hs_commit(s); }
extern "C" __attribute__((visibility("default"))) int hs_commit(shim_scene* s) {
    int rc = s->sb.flatten(s->flat);
    if (rc < 0) return set_err(rc, s->sb.err);
    s->has_media = false;
    for (const DevObject& o : s->flat.objects) if (o.flags & OBJ_MEDIUM) s->has_media = true;
    s->committed = true;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int hs_trace_closest(shim_scene* s, const float* rays, int64_t n, float t_min,
                                                                      float t_max, uint64_t seed, int32_t* prim, float* t,
                                                                      uint64_t* counters) {
    SceneView sv = view_of(s);
    uint64_t nodes = 0, prims = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* q = rays + i * 7;
        Ray r; r.o = mk3(q[0], q[1], q[2]); r.d = mk3(q[3], q[4], q[5]); r.time = q[6];
        Rng rng;
        rng_init(rng, (uint32_t)i, 0, seed);
        rng_key(rng, 0, STAGE_INTERSECT);
        TraceCounters tc; tc.nodes = 0; tc.prims = 0; tc.hrpp_tp = 0; tc.hrpp_fp = 0; tc.hrpp_none = 0;
        Hit h = hrpp_on(s) ? closest_hit<true, true>(sv, r, t_min, t_max, rng, &tc) : closest_hit<true, false>(sv, r, t_min, t_max, rng, &tc);
        hrpp_count(s, tc);
        nodes += tc.nodes; prims += tc.prims;
        prim[i] = hit_handle(sv, h);
        t[i] = h.obj < 0 ? INFINITY : h.t;
    }
    if (counters) { counters[0] = (uint64_t)n; counters[1] = nodes; counters[2] = prims; }
    return 0;
}

// closest hit of a one-Bvh world with closest_hit_solo, on the plain node layout or the signed one (SNode):
// the two must agree bit for bit, counters included
extern "C" __attribute__((visibility("default"))) int hs_trace_closest_solo(shim_scene* s, const float* rays, int64_t n, float t_min,
                                                                           float t_max, int signed_nodes, int32_t* prim, float* t,
                                                                           uint64_t* counters) {
    if (s->flat.objects.size() != 1 || s->flat.objects[0].kind != OBJ_BVH || s->flat.objects[0].flags != 0) return -1;
    if (signed_nodes && s->flat.snodes.empty()) s->flat.build_signed_nodes();
    SceneView sv = view_of(s);
    uint64_t nodes = 0, prims = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* q = rays + i * 7;
        Ray r; r.o = mk3(q[0], q[1], q[2]); r.d = mk3(q[3], q[4], q[5]); r.time = q[6];
        TraceCounters tc; tc.nodes = 0; tc.prims = 0; tc.hrpp_tp = 0; tc.hrpp_fp = 0; tc.hrpp_none = 0;
        Hit h = signed_nodes ? closest_hit_solo<true, -1, true>(sv, r, t_min, t_max, &tc) : closest_hit_solo<true, -1, false>(sv, r, t_min, t_max, &tc);
        nodes += tc.nodes; prims += tc.prims;
        prim[i] = hit_handle(sv, h);
        t[i] = h.obj < 0 ? INFINITY : h.t;
    }
    if (counters) { counters[0] = (uint64_t)n; counters[1] = nodes; counters[2] = prims; }
    return 0;
}

// closest hit of a world with ONE Bvh object among plain objects, the Bvh walked on its quantised nodes (QNode, what
// wf_bvh1_walk reads): must equal the walk over the exact boxes in id and t; returns the number of QNodes or -1
extern "C" __attribute__((visibility("default"))) int hs_trace_closest_q(shim_scene* s, const float* rays, int64_t n, float t_min,
                                                                        float t_max, int32_t* prim, float* t, uint64_t* counters) {
    int idx = -1, n_bvh = 0;
    for (size_t i = 0; i < s->flat.objects.size(); ++i) if (s->flat.objects[i].kind == OBJ_BVH) { ++n_bvh; idx = (int)i; }
    if (n_bvh != 1) return -1;
    if (s->flat.q_object != idx && !s->flat.build_quantized_nodes(idx)) return -1;
    SceneView sv = view_of(s);
    uint64_t nodes = 0, prims = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* q = rays + i * 7;
        Ray r; r.o = mk3(q[0], q[1], q[2]); r.d = mk3(q[3], q[4], q[5]); r.time = q[6];
        Rng rng;
        rng_init(rng, (uint32_t)i, 0, 0);
        rng_key(rng, 0, STAGE_INTERSECT);
        TraceCounters tc; tc.nodes = 0; tc.prims = 0; tc.hrpp_tp = 0; tc.hrpp_fp = 0; tc.hrpp_none = 0;
        Hit h = closest_hit<true, false, true, true>(sv, r, t_min, t_max, rng, &tc);
        nodes += tc.nodes; prims += tc.prims;
        prim[i] = hit_handle(sv, h);
        t[i] = h.obj < 0 ? INFINITY : h.t;
    }
    if (counters) { counters[0] = (uint64_t)n; counters[1] = nodes; counters[2] = prims; }
    return (int)s->flat.qnodes.size();
}
// the QNode array (8 words per node) for inspection
extern "C" __attribute__((visibility("default"))) int hs_qnodes(shim_scene* s, uint32_t* out, int max_nodes) {
    int n = (int)s->flat.qnodes.size();
    if (out) memcpy(out, s->flat.qnodes.data(), sizeof(QNode) * (size_t)(n < max_nodes ? n : max_nodes));
    return n;
}

// the wavefront's per-path arithmetic, executed path by path (same order of operations as
// wf_generate / wf_extend / wf_shade)
extern "C" __attribute__((visibility("default"))) int hs_sample_radiance(shim_scene* s, const shim_camera* cam,
                                                                        const shim_render_params* p, const int32_t* xys, int64_t n,
                                                                        float* out_rgb, uint64_t* rays_out) {
    SceneView sv = view_of(s);
    CameraPod c;
    camera_new(cam->look_from, cam->look_at, cam->view_up, cam->vertical_fov, cam->aspect_ratio, cam->aperture, cam->focus_dist,
               cam->time_start, cam->time_end, c);
    uint64_t rays = 0;
    for (int64_t i = 0; i < n; ++i) {
        int x = xys[i * 3], y = xys[i * 3 + 1];
        uint32_t sample = (uint32_t)xys[i * 3 + 2], pixel = (uint32_t)(y * p->width + x);
        Rng rng;
        rng_init(rng, pixel, sample, p->seed);
        rng_key(rng, 0, STAGE_CAMERA);
        Ray r = camera_sample(c, x, y, p->width, p->height, rng);
        f3 thr = mk3(1, 1, 1), L = mk3(0, 0, 0);
        for (int bounce = 0; bounce < p->max_depth; ++bounce) {
            rng_init(rng, pixel, sample, p->seed);
            rng_key(rng, (uint32_t)bounce, STAGE_INTERSECT);
            ++rays;
            if (ray_has_nan(r)) break;  // as wf_extend: a NaN ray ends the path
            TraceCounters tc; tc.nodes = 0; tc.prims = 0; tc.hrpp_tp = 0; tc.hrpp_fp = 0; tc.hrpp_none = 0;
            Hit h = hrpp_on(s) ? closest_hit<false, true>(sv, r, 0.001f, INFINITY, rng, &tc) : closest_hit<false, false>(sv, r, 0.001f, INFINITY, rng, &tc);
            hrpp_count(s, tc);
            if (h.obj < 0) { L = L + mk3(thr.x * p->background[0], thr.y * p->background[1], thr.z * p->background[2]); break; }
            int mat = hit_material(sv, h);
            int kind = mat_kind(sv, mat);
            HitRec rec;
            reconstruct_hit(sv, r, h, mat_needs_uv(sv, mat), rec);
            if (kind == MAT_DIFFUSE_LIGHT) { L = L + thr * mat_emit(sv, mat, rec); break; }
            rng_init(rng, pixel, sample, p->seed);
            rng_key(rng, (uint32_t)bounce, STAGE_SCATTER);
            f3 att; Ray out;
            if (!mat_scatter(sv, kind, mat, r, rec, rng, att, out)) break;
            thr = thr * att;
            r = out;
        }
        out_rgb[i * 3] = L.x; out_rgb[i * 3 + 1] = L.y; out_rgb[i * 3 + 2] = L.z;
    }
    if (rays_out) *rays_out = rays;
    return 0;
}

extern "C" __attribute__((visibility("default"))) void hs_texture_value(shim_scene* s, int tex, float u, float v, const float* p,
                                                                       float* out) {
    SceneView sv = s->flat.view();
    f3 c = tex_value(sv, tex, u, v, mk3(p[0], p[1], p[2]));
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

extern "C" __attribute__((visibility("default"))) void hs_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out[0], out[1], out[2], out[3]);
}
