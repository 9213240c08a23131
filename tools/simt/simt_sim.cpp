// simt_sim.cpp — ANALYSIS TOOL (not part of the package, never loaded by it).
//
// Replays the wavefront of a one-Bvh world on the CPU with the product's own __host__ __device__ arithmetic
// (csrc/shim_device.h, linked with tests/hostsim) and models how a 32-lane warp would execute the closest-hit
// walk of every queue entry: which lanes are busy in every node step and every primitive test of the while-while
// loop.  It answers, without GPU time, what SIMT efficiency a scheduling policy can reach on the real ray
// population of a render (queue order included):
//   policy 0  one ray per lane, the warp runs until its longest walk ends (wf_trace_solo as built)
//   policy 1  a per-warp pool of K*32 rays in shared memory, idle lanes take the next ray of the pool once at
//             least R lanes are idle (walk phase of a shade -> pool -> walk -> append kernel)
// Costs are issue slots per warp instruction group, taken from the SASS of the round-1 kernel.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "../../raytracinginoneweekendinrust_b200/csrc/shim_device.h"
#include "../../raytracinginoneweekendinrust_b200/csrc/shim_internal.h"

using namespace shim;

namespace {

// the walk of bvh_closest (shim_device.h) with a token per step: 'N' node step, 'P' primitive test
bool walk_tokens(const SceneView& sv, int start, const RayCtx& c, float t_min, float t_max, BvhBest& best, std::vector<char>& tok) {
    best.t = t_max; best.prim = 0; best.face = 0; best.any = false;
    float t_cull = t_max;
    int stack[SHIM_BVH_STACK];
    int sp = 0, cur = start;
    for (;;) {
        while (cur >= 0 && cur != SHIM_STACK_END) {
            const DevNode& n = sv.nodes[cur];
            tok.push_back('N');
            float tl, tr;
            bool hl = slab(n.a.x, n.a.y, n.a.z, n.a.w, n.b.x, n.b.y, c, t_min, t_cull, tl);
            bool hr = slab(n.b.z, n.b.w, n.c.x, n.c.y, n.c.z, n.c.w, c, t_min, t_cull, tr);
            hr = hr && n.d.y != CHILD_NONE;
            if (hl && hr) { bool swap = tr < tl; cur = swap ? n.d.y : n.d.x; if (sp < SHIM_BVH_STACK) stack[sp++] = swap ? n.d.x : n.d.y; }
            else if (hl) cur = n.d.x;
            else if (hr) cur = n.d.y;
            else cur = sp > 0 ? stack[--sp] : SHIM_STACK_END;
        }
        if (cur == SHIM_STACK_END) break;
        uint32_t ref = ~(uint32_t)cur;
        float t; int face = 0;
        tok.push_back('P');
        if (hit_prim(sv, ref, c, t_min, t_cull, t, face) && !(t > best.t)) {
            bool take = !best.any || t < best.t;
            if (!take) take = tie_goes_to_candidate(sv, c, t_min, ref, best.prim, t);
            if (take) { best.t = t; best.prim = ref; best.face = face; best.any = true; t_cull = t + fabsf(t) * 3.8146973e-06f; }
        }
        cur = sp > 0 ? stack[--sp] : SHIM_STACK_END;
    }
    return best.any;
}

struct Entry { Ray r; f3 thr; uint32_t pixel, sample; int bounce; Hit h; int mat; };
struct Cost { double node, prim, refill, setup; };

// lanes execute their token strings in lockstep: inner loop while any lane's next token is 'N', then one 'P' round
struct WarpSim {
    double slots = 0, useful = 0;   // issue slots spent by the warp x 32 / slots in which a lane did work
    double bound_max = 0;           // sum over warps of the longest lane (x 32): what walk-length variance alone costs
    // mode 0: while-while (node steps until no lane wants one, then one round of primitive tests)
    // mode 1: each round runs the step kind that more lanes are waiting for (primitive tests are postponed until
    //         `thr` lanes want one or no lane wants a node step)
    // mode 2: if-if (every round runs a node step and a primitive test for whoever wants one)
    void run(const std::vector<const std::vector<char>*>& rays, int K, int R, const Cost& c, int mode = 0, int thr = 16) {
        // rays: the warp's pool in order; K = 0: policy 0 (rays.size() <= 32)
        const std::vector<char>* cur[32]; size_t pos[32];
        size_t next = 0;
        for (int l = 0; l < 32; ++l) { cur[l] = nullptr; pos[l] = 0; }
        auto idle = [&](int l) { return cur[l] == nullptr || pos[l] >= cur[l]->size(); };
        if (!K) {
            double mx = 0;
            for (auto* r : rays) { double a = c.setup; for (char t : *r) a += t == 'N' ? c.node : c.prim; if (a > mx) mx = a; }
            bound_max += 32 * mx;
        }
        for (;;) {
            int n_idle = 0;
            for (int l = 0; l < 32; ++l) n_idle += idle(l);
            if (next < rays.size() && (n_idle >= R || n_idle == 32)) {   // refill
                int took = 0;
                for (int l = 0; l < 32 && next < rays.size(); ++l) if (idle(l)) { cur[l] = rays[next++]; pos[l] = 0; ++took; }
                slots += 32 * (K ? c.refill : c.setup); useful += took * (K ? c.refill : c.setup);
                continue;
            }
            if (n_idle == 32) break;
            if (mode == 1 || mode == 2) {
                int wn = 0, wp = 0;
                for (int l = 0; l < 32; ++l) if (!idle(l)) { if ((*cur[l])[pos[l]] == 'N') ++wn; else ++wp; }
                const bool do_p = mode == 2 ? wp > 0 : (wp >= thr || wn == 0);
                const bool do_n = mode == 2 ? wn > 0 : !do_p;
                if (do_n) { for (int l = 0; l < 32; ++l) if (!idle(l) && (*cur[l])[pos[l]] == 'N') ++pos[l]; slots += 32 * c.node; useful += wn * c.node; }
                if (do_p) { for (int l = 0; l < 32; ++l) if (!idle(l) && (*cur[l])[pos[l]] == 'P' && !(mode == 2 && do_n && false)) ++pos[l]; slots += 32 * c.prim; useful += wp * c.prim; }
                continue;
            }
            // inner while: node steps
            for (;;) {
                int nn = 0;
                for (int l = 0; l < 32; ++l) if (!idle(l) && (*cur[l])[pos[l]] == 'N') { ++pos[l]; ++nn; }
                if (!nn) break;
                slots += 32 * c.node; useful += nn * c.node;
            }
            int np = 0;
            for (int l = 0; l < 32; ++l) if (!idle(l) && (*cur[l])[pos[l]] == 'P') { ++pos[l]; ++np; }
            if (np) { slots += 32 * c.prim; useful += np * c.prim; }
        }
    }
};

}  // namespace

// out: per policy p and bounce b (0..max_b-1): [rays, slots, useful, longest-lane bound]; policies: (K, R, mode, thr) rows
extern "C" __attribute__((visibility("default"))) int simt_sim(shim_scene* s, const shim_camera* cam, const shim_render_params* p,
                                                                const int32_t* xys, int64_t n, const int* policies, int n_pol,
                                                                const double* cost4, int max_b, double* out) {
    SceneView sv = s->flat.view();
    CameraPod cp;
    camera_new(cam->look_from, cam->look_at, cam->view_up, cam->vertical_fov, cam->aspect_ratio, cam->aperture, cam->focus_dist,
               cam->time_start, cam->time_end, cp);
    Cost cost{cost4[0], cost4[1], cost4[2], cost4[3]};
    memset(out, 0, sizeof(double) * 4 * n_pol * max_b);
    // queue of the current iteration: camera rays first (tile order), then per material kind
    std::vector<Entry> q[MAT_KINDS + 1];   // [0] = new camera rays, [1 + kind] = material queues
    for (int64_t i = 0; i < n; ++i) {
        Entry e;
        int x = xys[i * 3], y = xys[i * 3 + 1];
        e.sample = (uint32_t)xys[i * 3 + 2]; e.pixel = (uint32_t)(y * p->width + x); e.bounce = 0;
        Rng rng; rng_init(rng, e.pixel, e.sample, p->seed); rng_key(rng, 0, STAGE_CAMERA);
        e.r = camera_sample(cp, x, y, p->width, p->height, rng);
        e.thr = mk3(1, 1, 1);
        q[0].push_back(e);
    }
    for (int iter = 0; iter < max_b; ++iter) {
        std::vector<Entry> nq[MAT_KINDS + 1];
        size_t total = 0;
        for (int seg = 0; seg <= MAT_KINDS; ++seg) {
            // shade the segment's entries (except camera rays) -> rays to walk, in order
            std::vector<Entry> rays;
            for (Entry& e : q[seg]) {
                if (seg > 0) {
                    HitRec rec;
                    reconstruct_hit(sv, e.r, e.h, mat_needs_uv(sv, e.mat), rec);
                    const int kind = seg - 1;
                    if (kind == MAT_DIFFUSE_LIGHT) continue;
                    Rng rng; rng_init(rng, e.pixel, e.sample, p->seed); rng_key(rng, (uint32_t)e.bounce, STAGE_SCATTER);
                    f3 att; Ray o;
                    if (!mat_scatter(sv, kind, e.mat, e.r, rec, rng, att, o)) continue;
                    if (e.bounce + 1 >= p->max_depth) continue;
                    e.r = o; e.thr = e.thr * att; e.bounce += 1;
                }
                if (ray_has_nan(e.r)) continue;
                rays.push_back(e);
            }
            total += rays.size();
            // token strings + hits
            std::vector<std::vector<char>> tok(rays.size());
            for (size_t i = 0; i < rays.size(); ++i) {
                RayCtx c; make_ctx(c, rays[i].r);
                BvhBest best;
                if (walk_tokens(sv, sv.objects[0].ref, c, 0.001f, SHIM_INF, best, tok[i])) {
                    Entry e = rays[i];
                    e.h.t = best.t; e.h.obj = 0; e.h.prim = best.prim; e.h.face = best.face;
                    const int mw = hit_material_word(sv, e.h);
                    e.mat = mat_word_index(mw);
                    nq[1 + mat_word_kind(mw)].push_back(e);
                }
            }
            for (int pi = 0; pi < n_pol; ++pi) {
                const int K = policies[4 * pi], R = policies[4 * pi + 1], mode = policies[4 * pi + 2], thr = policies[4 * pi + 3];
                const size_t per = K ? (size_t)K * 32 : 32;
                WarpSim ws;
                for (size_t b0 = 0; b0 < rays.size(); b0 += per) {
                    std::vector<const std::vector<char>*> pool;
                    for (size_t i = b0; i < rays.size() && i < b0 + per; ++i) pool.push_back(&tok[i]);
                    ws.run(pool, K, K ? R : 32, cost, mode, thr);
                }
                double* o = out + 4 * (pi * max_b + iter);
                o[0] += (double)rays.size(); o[1] += ws.slots; o[2] += ws.useful; o[3] += ws.bound_max;
            }
        }
        // experiment (SIMT_SORT_WINDOW=N): the entries of each material queue sorted by the Morton key of their hit point
        // inside windows of N consecutive entries (what a block could do in shared memory before it appends)
        if (const char* sw = getenv("SIMT_SORT_WINDOW")) {
            const size_t win = (size_t)atol(sw);
            auto key = [](const Entry& e) {
                const f3 pt = e.r.o + e.h.t * e.r.d;
                auto q10 = [](float v) { float u = (v + 16.0f) / 32.0f; u = u < 0 ? 0 : (u > 0.999f ? 0.999f : u); return (uint32_t)(u * 1024.0f); };
                auto spread = [](uint32_t x) { x &= 1023u; x = (x | (x << 16)) & 0x30000ffu; x = (x | (x << 8)) & 0x300f00fu; x = (x | (x << 4)) & 0x30c30c3u; x = (x | (x << 2)) & 0x9249249u; return x; };
                return spread(q10(pt.x)) | (spread(q10(pt.y * 8.0f - 12.0f)) << 1) | (spread(q10(pt.z)) << 2);
            };
            for (int seg = 1; seg <= MAT_KINDS; ++seg)
                for (size_t b0 = 0; win && b0 < nq[seg].size(); b0 += win) {
                    auto first = nq[seg].begin() + b0, last = nq[seg].begin() + std::min(nq[seg].size(), b0 + win);
                    std::stable_sort(first, last, [&](const Entry& a, const Entry& b) { return key(a) < key(b); });
                }
        }
        for (int seg = 0; seg <= MAT_KINDS; ++seg) q[seg].swap(nq[seg]);
        q[0].clear();
        if (total == 0) break;
    }
    return 0;
}
