// shim_scene.cpp — host builder, bvh.rs-style tree build and the flattener.
#include "shim_scene.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <functional>
#include <thread>

namespace shim {

static const float kEps = 1.1920929e-07f;  // f32::EPSILON
static const float kPi = 3.14159265358979323846f;

static Box box_union(const Box& a, const Box& b) {  // Aabb::union, aabb.rs:43-62
    Box r;
    for (int i = 0; i < 3; ++i) { r.mn[i] = std::fmin(a.mn[i], b.mn[i]); r.mx[i] = std::fmax(a.mx[i], b.mx[i]); }
    return r;
}

bool SceneBuilder::bounding_box(int hid, float t0, float t1, Box& out) const {
    const HostHittable& h = hittables[hid];
    const float* p = h.p;
    switch (h.kind) {
    case H_SPHERE:  // sphere.rs:105-109
        for (int i = 0; i < 3; ++i) { out.mn[i] = p[i] - p[3]; out.mx[i] = p[i] + p[3]; }
        return true;
    case H_MSPHERE: {  // moving_sphere.rs:86-93; the true union of the start and end boxes
        // (the reference's end box reuses center(time_0) for its min, which equals this box
        // whenever the motion is non-negative on every axis, as in all of its scenes)
        Box a, b;
        for (int i = 0; i < 3; ++i) {
            float ca = p[i] + ((t0 - p[6]) / (p[7] - p[6])) * (p[3 + i] - p[i]);
            float cb = p[i] + ((t1 - p[6]) / (p[7] - p[6])) * (p[3 + i] - p[i]);
            a.mn[i] = ca - p[8]; a.mx[i] = ca + p[8];
            b.mn[i] = cb - p[8]; b.mx[i] = cb + p[8];
        }
        out = box_union(a, b);
        return true;
    }
    case H_RECT: {  // rectangle.rs:67-73 / 129-135 / 191-197
        int ia = h.axis == 0 ? 1 : 0, ib = h.axis == 2 ? 1 : 2;
        out.mn[ia] = p[0]; out.mx[ia] = p[1];
        out.mn[ib] = p[2]; out.mx[ib] = p[3];
        out.mn[h.axis] = p[4] - kEps; out.mx[h.axis] = p[4] + kEps;
        return true;
    }
    case H_TRI:  // triangle.rs:94-107
        for (int i = 0; i < 3; ++i) {
            out.mn[i] = std::fmin(p[i], std::fmin(p[3 + i], p[6 + i])) - kEps;
            out.mx[i] = std::fmax(p[i], std::fmax(p[3 + i], p[6 + i])) + kEps;
        }
        return true;
    case H_CUBE:  // cube.rs:95-97
        for (int i = 0; i < 3; ++i) { out.mn[i] = p[i]; out.mx[i] = p[3 + i]; }
        return true;
    default:
        return false;
    }
}

static int total_cmp(float a, float b) {  // f32::total_cmp
    int32_t l, r;
    memcpy(&l, &a, 4); memcpy(&r, &b, 4);
    l ^= (int32_t)(((uint32_t)(l >> 31)) >> 1);
    r ^= (int32_t)(((uint32_t)(r >> 31)) >> 1);
    return l < r ? -1 : (l > r ? 1 : 0);
}
static uint64_t splitmix64(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int bvh_height(const std::vector<HostBvhNode>& nodes, int i) {  // BvhNode::max_depth, bvh.rs:335-350
    int l = nodes[i].left >= 0 ? bvh_height(nodes, nodes[i].left) : 0;
    int r = nodes[i].right >= 0 ? bvh_height(nodes, nodes[i].right) : 0;
    return (l > r ? l : r) + 1;
}

static int build_device_tree(const SceneBuilder& sb, const HostHittable& b, std::vector<HostBvhNode>& out);
// The device tree belongs to the BVH's construction (Bvh::new, bvh.rs:46-62, is scene-build time in the reference
// too), so shim_commit only lays it out.
static void attach_device_tree(const SceneBuilder& sb, HostHittable& bvh) {
    bvh.dev_root = build_device_tree(sb, bvh, bvh.dev_nodes);
    if (bvh.dev_root >= 0) bvh.dev_height = bvh_height(bvh.dev_nodes, bvh.dev_root);
    else bvh.dev_nodes.clear();
}

int SceneBuilder::build_bvh(int list, float t0, float t1, uint64_t axis_seed, bool predictor) {
    if (!ok_hit(list) || hittables[list].kind != H_LIST) { err = "shim_bvh: not a list"; return SHIM_ERR_INVALID_; }
    std::vector<int> objs = hittables[list].items;
    if (objs.empty()) { err = "shim_bvh: empty list"; return SHIM_ERR_INVALID_; }
    // cache the comparator boxes (bounding_box(0,0), bvh.rs:420-428) and the node boxes (t0,t1)
    std::unordered_map<int, Box> cmp_box, node_box;
    for (int o : objs) {
        Box a, b;
        if (!bounding_box(o, 0.0f, 0.0f, a) || !bounding_box(o, t0, t1, b)) {
            err = "shim_bvh: a BVH may only contain spheres, moving spheres, rects, triangles and cubes";
            return SHIM_ERR_UNSUPPORTED_;
        }
        cmp_box[o] = a; node_box[o] = b;
    }
    HostHittable bvh;
    bvh.kind = H_BVH; bvh.child = list; bvh.t0 = t0; bvh.t1 = t1; bvh.predictor = predictor;
    bvh.nodes.reserve(objs.size() * 2 + 1);
    uint64_t state = axis_seed;
    std::function<int(int*, size_t)> rec = [&](int* o, size_t n) -> int {
        int axis = (int)(splitmix64(state) % 3ull);  // stands in for rng.gen_range(0..=2), bvh.rs:255-257
        auto less = [&](int a, int b) { return total_cmp(cmp_box[a].mn[axis], cmp_box[b].mn[axis]) < 0; };
        HostBvhNode node;
        node.parent = -1;
        if (n == 1) { node.left = ~o[0]; node.right = ~o[0]; }
        else if (n == 2) {
            if (less(o[0], o[1])) { node.left = ~o[0]; node.right = ~o[1]; }
            else { node.left = ~o[1]; node.right = ~o[0]; }
        } else {
            std::stable_sort(o, o + n, less);
            size_t mid = n / 2;
            node.left = rec(o, mid);
            node.right = rec(o + mid, n - mid);
        }
        Box lb = node.left >= 0 ? bvh.nodes[node.left].box : node_box[~node.left];
        Box rb = node.right >= 0 ? bvh.nodes[node.right].box : node_box[~node.right];
        node.box = box_union(lb, rb);
        int idx = (int)bvh.nodes.size();
        if (node.left >= 0) bvh.nodes[node.left].parent = idx;
        if (node.right >= 0) bvh.nodes[node.right].parent = idx;
        bvh.nodes.push_back(node);
        return idx;
    };
    bvh.root = rec(objs.data(), objs.size());
    bvh.height = bvh_height(bvh.nodes, bvh.root);
    attach_device_tree(*this, bvh);
    return add_hittable(bvh);
}

int SceneBuilder::bvh_from_nodes(int n, const int32_t* left, const int32_t* right, int root, float t0, float t1, bool predictor) {
    if (n <= 0 || root < 0 || root >= n || !left || !right) { err = "shim_bvh_from_nodes: bad arguments"; return SHIM_ERR_INVALID_; }
    HostHittable bvh;
    bvh.kind = H_BVH; bvh.t0 = t0; bvh.t1 = t1; bvh.predictor = predictor; bvh.root = root;
    bvh.nodes.resize(n);
    std::vector<int> state(n, 0);
    for (int i = 0; i < n; ++i) { bvh.nodes[i].left = left[i]; bvh.nodes[i].right = right[i]; bvh.nodes[i].parent = -1; }
    // boxes bottom-up (iterative post-order); also validates that the graph is a tree
    std::vector<int> stack{root};
    while (!stack.empty()) {
        int i = stack.back();
        HostBvhNode& nd = bvh.nodes[i];
        if (state[i] == 0) {
            state[i] = 1;
            for (int c : {nd.left, nd.right}) {
                if (c >= 0) {
                    if (c >= n || state[c] != 0) { err = "shim_bvh_from_nodes: not a tree"; return SHIM_ERR_INVALID_; }
                    bvh.nodes[c].parent = i;
                    stack.push_back(c);
                } else if (!ok_hit(~c)) { err = "shim_bvh_from_nodes: bad primitive id"; return SHIM_ERR_INVALID_; }
            }
        } else {
            stack.pop_back();
            Box b[2];
            int k = 0;
            for (int c : {nd.left, nd.right}) {
                if (c >= 0) b[k] = bvh.nodes[c].box;
                else if (!bounding_box(~c, t0, t1, b[k])) { err = "shim_bvh_from_nodes: unsupported BVH member"; return SHIM_ERR_UNSUPPORTED_; }
                ++k;
            }
            nd.box = box_union(b[0], b[1]);
            state[i] = 2;
        }
    }
    bvh.height = bvh_height(bvh.nodes, root);
    attach_device_tree(*this, bvh);
    return add_hittable(bvh);
}

// ---------------------------------------------------------------------------- device tree (binned SAH)
// The closest hit does not depend on the tree, so the tree the kernels walk is
// rebuilt from the BVH's primitive list with a binned surface-area heuristic instead of bvh.rs's
// random-axis median split (19 -> ~8 node visits per ray on the Book-1 scene).  The reference
// topology stays in HostHittable::nodes (shim_bvh_nodes) and can be selected for the device too.
namespace {
struct SahBuilder {
    const std::vector<int>& prims;            // hittable ids
    const std::vector<Box>& boxes;            // box of prims[i]
    std::vector<HostBvhNode>& out;
    std::vector<int> idx;
    static float area(const Box& b) {
        float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
        if (!(dx >= 0) || !(dy >= 0) || !(dz >= 0)) return 0.0f;
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
    static Box empty() { Box b; for (int k = 0; k < 3; ++k) { b.mn[k] = INFINITY; b.mx[k] = -INFINITY; } return b; }
    // plain compares (std::fmin / std::fmax are libm calls and dominated the build time)
    static Box join(const Box& a, const Box& b) {
        Box r;
        for (int k = 0; k < 3; ++k) { r.mn[k] = a.mn[k] < b.mn[k] ? a.mn[k] : b.mn[k]; r.mx[k] = a.mx[k] > b.mx[k] ? a.mx[k] : b.mx[k]; }
        return r;
    }
    static int ceil_log2(size_t n) { int l = 0; while (((size_t)1 << l) < n) ++l; return l; }

    // returns a child reference: >= 0 node index, < 0 ~hittable id
    int build(size_t lo, size_t hi, int depth, Box& box_out) {
        size_t n = hi - lo;
        if (n == 1) { box_out = boxes[idx[lo]]; return ~prims[idx[lo]]; }
        Box bounds = empty(), cb = empty();
        for (size_t i = lo; i < hi; ++i) {
            const Box& b = boxes[idx[i]];
            bounds = join(bounds, b);
            for (int k = 0; k < 3; ++k) {
                float c = 0.5f * (b.mn[k] + b.mx[k]);
                cb.mn[k] = c < cb.mn[k] ? c : cb.mn[k]; cb.mx[k] = c > cb.mx[k] ? c : cb.mx[k];
            }
        }
        size_t mid = lo + n / 2;
        bool median = depth + ceil_log2(n) >= SHIM_MAX_BVH_HEIGHT - 2;  // keep the height inside the traversal stack
        int best_axis = -1; int best_bin = -1;
        const int NBMAX = 32;
        const int NB = n <= 8 ? 8 : (n <= 64 ? 16 : NBMAX);   // small nodes dominate the node count: fewer bins there
        if (!median && n > 2) {
            float best_cost = INFINITY;
            for (int axis = 0; axis < 3; ++axis) {
                float ext = cb.mx[axis] - cb.mn[axis];
                if (!(ext > 0.0f)) continue;
                Box bb[NBMAX]; int cnt[NBMAX];
                for (int b = 0; b < NB; ++b) { bb[b] = empty(); cnt[b] = 0; }
                float scale = (float)NB / ext;
                for (size_t i = lo; i < hi; ++i) {
                    const Box& bx = boxes[idx[i]];
                    int b = (int)((0.5f * (bx.mn[axis] + bx.mx[axis]) - cb.mn[axis]) * scale);
                    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                    bb[b] = join(bb[b], bx); cnt[b]++;
                }
                float right_area[NBMAX]; int right_cnt[NBMAX];
                Box acc = empty(); int c = 0;
                for (int b = NB - 1; b > 0; --b) { acc = join(acc, bb[b]); c += cnt[b]; right_area[b] = area(acc); right_cnt[b] = c; }
                acc = empty(); c = 0;
                for (int b = 0; b < NB - 1; ++b) {
                    acc = join(acc, bb[b]); c += cnt[b];
                    if (c == 0 || right_cnt[b + 1] == 0) continue;
                    float cost = area(acc) * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
                }
            }
        }
        if (best_axis >= 0) {
            float ext = cb.mx[best_axis] - cb.mn[best_axis];
            float scale = (float)NB / ext;
            auto it = std::partition(idx.begin() + lo, idx.begin() + hi, [&](int i) {
                const Box& bx = boxes[i];
                int b = (int)((0.5f * (bx.mn[best_axis] + bx.mx[best_axis]) - cb.mn[best_axis]) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                return b <= best_bin;
            });
            mid = (size_t)(it - idx.begin());
            if (mid == lo || mid == hi) mid = lo + n / 2;
        } else if (n > 2) {
            int axis = 0;
            for (int k = 1; k < 3; ++k) if (cb.mx[k] - cb.mn[k] > cb.mx[axis] - cb.mn[axis]) axis = k;
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
                return boxes[a].mn[axis] + boxes[a].mx[axis] < boxes[b].mn[axis] + boxes[b].mx[axis];
            });
        }
        HostBvhNode node;
        node.parent = -1;
        Box lb, rb;
        node.left = build(lo, mid, depth + 1, lb);
        node.right = build(mid, hi, depth + 1, rb);
        node.box = join(lb, rb);
        int me = (int)out.size();
        if (node.left >= 0) out[node.left].parent = me;
        if (node.right >= 0) out[node.right].parent = me;
        out.push_back(node);
        box_out = node.box;
        return me;
    }
};
}  // namespace

// builds the device tree of a BVH hittable; returns the root node index inside `out`
static int build_device_tree(const SceneBuilder& sb, const HostHittable& b, std::vector<HostBvhNode>& out) {
    // the primitive set = the leaves of the recorded tree, in left-to-right order
    std::vector<int> prims;
    std::vector<int> stack{b.root};
    while (!stack.empty()) {
        int i = stack.back(); stack.pop_back();
        if (i < 0) { int h = ~i; if (prims.empty() || prims.back() != h) prims.push_back(h); continue; }
        const HostBvhNode& n = b.nodes[i];
        stack.push_back(n.right);
        stack.push_back(n.left);
    }
    std::vector<Box> boxes(prims.size());
    for (size_t i = 0; i < prims.size(); ++i)
        if (!sb.ok_hit(prims[i]) || !sb.bounding_box(prims[i], b.t0, b.t1, boxes[i])) return -1;
    out.clear();
    out.reserve(prims.size());
    SahBuilder sah{prims, boxes, out, {}};
    sah.idx.resize(prims.size());
    for (size_t i = 0; i < prims.size(); ++i) sah.idx[i] = (int)i;
    Box bx;
    int root = sah.build(0, prims.size(), 0, bx);
    if (root < 0) {  // a single primitive: a node with one child, like bvh.rs's len == 1 case
        HostBvhNode node; node.parent = -1; node.left = root; node.right = root; node.box = bx;
        out.push_back(node);
        root = 0;
    }
    return root;
}

// ---------------------------------------------------------------------------- Perlin tables
static void perlin_table(uint32_t seed, uint8_t* perm) {
    for (int i = 0; i < 256; ++i) perm[i] = (uint8_t)i;
    uint64_t s = 0x5851F42D4C957F2Dull ^ (uint64_t)seed;
    for (int i = 255; i > 0; --i) {
        uint32_t j = (uint32_t)(splitmix64(s) % (uint64_t)(i + 1));
        std::swap(perm[i], perm[j]);
    }
}
void marble_tables(uint32_t seed, std::vector<uint8_t>& out) {
    size_t base = out.size();
    out.resize(base + 19 * 256);
    uint8_t* t = out.data() + base;
    perlin_table(seed, t);  // Turbulence source = Perlin::new(seed), marble.rs:14-19
    for (int f = 0; f < 3; ++f)          // x/y/z distortion Fbm seeded 0,1,2
        for (int o = 0; o < 6; ++o) perlin_table((uint32_t)(f + o), t + 256 * (1 + f * 6 + o));
}

// ---------------------------------------------------------------------------- flatten
namespace {
struct Flattener {
    SceneBuilder& sb;
    FlatScene& fs;
    std::vector<uint32_t> prim_of;               // hittable id -> prim_ref (NO_PRIM: not flattened yet); dense: ids are small
    static constexpr uint32_t NO_PRIM = 0xffffffffu;
    std::unordered_map<int, int> bvh_base;       // hittable id -> first node index
    std::unordered_map<int, int> bvh_pred;       // hittable id -> predictor index
    std::unordered_map<int, int> bvh_root, bvh_nodes;
    int status = 0;

    struct Xf { bool translate = false, rotate = false; float d[3] = {0, 0, 0}; float s = 0, c = 1; int medium = -1; };

    int fail(int code, const std::string& m) { sb.err = m; status = code; return code; }
    // material reference as the device stores it: index | kind << 28 (shim_device.h: mat_word_index / mat_word_kind)
    // The kind field is the shading class (shim_types.h): a Lambertian whose texture tree holds a marble goes to the
    // class of its own - except in worlds that are one plain Bvh, whose pipeline keeps the five kinds.
    bool texture_is_slow(int t, int depth = 0) const {
        if (t < 0 || (size_t)t >= sb.textures.size() || depth > 8) return false;
        const HostTexture& x = sb.textures[t];
        if (x.kind == TEX_MARBLE) return true;
        return x.kind == TEX_CHECKER && (texture_is_slow(x.even, depth + 1) || texture_is_slow(x.odd, depth + 1));
    }
    int material_word(int mi) {
        if (mi < 0 || (size_t)mi >= sb.materials.size()) return mi;
        int cls = sb.materials[mi].kind;
        const bool one_plain_bvh = sb.world.size() == 1 && sb.ok_hit(sb.world[0]) && sb.hittables[sb.world[0]].kind == H_BVH;
        if (cls == MAT_LAMBERTIAN && !one_plain_bvh && texture_is_slow(sb.materials[mi].tex)) { cls = MQ_SLOW_LAMBERTIAN; fs.slow_lambertians = true; }
        return (int)((uint32_t)mi | ((uint32_t)cls << 28));
    }

    uint32_t add_prim(int hid) {
        if (prim_of.size() < sb.hittables.size()) prim_of.resize(sb.hittables.size(), NO_PRIM);
        if (prim_of[hid] != NO_PRIM) return prim_of[hid];
        const HostHittable& h = sb.hittables[hid];
        const float* p = h.p;
        uint32_t ref = 0;
        auto bits = [](int i) { float f; memcpy(&f, &i, 4); return f; };
        const int mat_word = material_word(h.material);
        switch (h.kind) {
        case H_SPHERE: {
            ref = prim_ref(PT_SPHERE, (uint32_t)fs.sph_s.size());
            double r = (double)p[3];
            fs.sph.push_back((double)p[0]); fs.sph.push_back((double)p[1]); fs.sph.push_back((double)p[2]); fs.sph.push_back(r * r);
            fs.sph_s.push_back(f4{p[0], p[1], p[2], p[3]});
            fs.sph_mat.push_back(mat_word);
            fs.handle[PT_SPHERE].push_back(hid);
            break;
        }
        case H_MSPHERE:
            ref = prim_ref(PT_MSPHERE, (uint32_t)fs.handle[PT_MSPHERE].size());
            fs.msph.push_back(f4{p[0], p[1], p[2], p[8]});
            fs.msph.push_back(f4{p[3], p[4], p[5], p[6]});
            fs.msph.push_back(f4{p[7], bits(mat_word), 0, 0});
            fs.handle[PT_MSPHERE].push_back(hid);
            break;
        case H_RECT:
            ref = prim_ref(PT_RECT, (uint32_t)fs.handle[PT_RECT].size());
            fs.rect.push_back(f4{p[0], p[1], p[2], p[3]});
            fs.rect.push_back(f4{p[4], bits(h.axis), bits(mat_word), 0});
            fs.handle[PT_RECT].push_back(hid);
            break;
        case H_TRI:
            ref = prim_ref(PT_TRI, (uint32_t)fs.handle[PT_TRI].size());
            fs.tri.push_back(f4{p[0], p[1], p[2], bits(mat_word)});
            fs.tri.push_back(f4{p[3] - p[0], p[4] - p[1], p[5] - p[2], 0});  // edge1 = vertex1 - vertex0, triangle.rs:44
            fs.tri.push_back(f4{p[6] - p[0], p[7] - p[1], p[8] - p[2], 0});  // edge2
            fs.handle[PT_TRI].push_back(hid);
            break;
        default:  // H_CUBE
            ref = prim_ref(PT_CUBE, (uint32_t)fs.handle[PT_CUBE].size());
            fs.cube.push_back(f4{p[0], p[1], p[2], bits(mat_word)});
            fs.cube.push_back(f4{p[3], p[4], p[5], 0});
            fs.handle[PT_CUBE].push_back(hid);
            break;
        }
        fs.rank[prim_type(ref)].push_back(0);
        fs.leaf[prim_type(ref)].push_back(-1);
        fs.sibling[prim_type(ref)].push_back(-1);
        prim_of[hid] = ref;
        return ref;
    }

    static bool is_prim(int kind) { return kind == H_SPHERE || kind == H_MSPHERE || kind == H_RECT || kind == H_TRI || kind == H_CUBE; }

    int add_bvh(int hid) {
        auto it = bvh_base.find(hid);
        if (it != bvh_base.end()) return it->second;
        const HostHittable& rec = sb.hittables[hid];
        // device tree: SAH rebuild by default, the recorded (bvh.rs) topology when asked for
        HostHittable sah_tree;
        if (!sb.device_reference_topology) {
            if (rec.dev_root < 0) { fail(SHIM_ERR_UNSUPPORTED_, "a BVH may only contain primitives and cubes"); return -1; }
            sah_tree.t0 = rec.t0; sah_tree.t1 = rec.t1; sah_tree.predictor = rec.predictor;
            sah_tree.nodes = rec.dev_nodes; sah_tree.root = rec.dev_root; sah_tree.height = rec.dev_height;
        }
        const HostHittable& b = sb.device_reference_topology ? rec : sah_tree;
        int base = (int)fs.nodes.size();
        fs.nodes.resize(base + b.nodes.size());
        // in-order walk: primitives are appended to the device arrays in left-to-right leaf order (neighbouring
        // leaves are neighbours in memory) and get their leaf rank (tie rule of bvh.rs:409-415) and leaf node
        if (b.height > SHIM_MAX_BVH_HEIGHT) { fail(SHIM_ERR_UNSUPPORTED_, "BVH taller than the traversal stack"); return -1; }
        // ties (bvh.rs:409-415) go to the primitive that is latest in the left-to-right leaf order of the RECORDED
        // bvh.rs tree, whichever tree the device walks
        std::vector<int> ref_rank(sb.hittables.size(), 0), ref_sibling(sb.hittables.size(), -1);   // sibling: the other primitive of a recorded two-primitive leaf
        for (const HostBvhNode& n : rec.nodes)
            if (n.left < 0 && n.right < 0 && n.left != n.right) { ref_sibling[~n.left] = ~n.right; ref_sibling[~n.right] = ~n.left; }
        {
            int r = 0;
            std::vector<int> st{rec.root};
            while (!st.empty()) {
                int i = st.back(); st.pop_back();
                if (i < 0) { ref_rank[~i] = r++; continue; }   // a duplicated single-object leaf keeps its later rank
                st.push_back(rec.nodes[i].right);
                st.push_back(rec.nodes[i].left);
            }
        }
        bool bad = false;
        std::vector<std::pair<uint32_t, int>> pending_siblings;   // (prim_ref, sibling hittable id): resolved once every primitive has its ref
        std::function<void(int)> walk = [&](int i) {
            const HostBvhNode& n = b.nodes[i];
            bool dup = n.left < 0 && n.right < 0 && n.left == n.right;
            auto leaf_child = [&](int c) {
                int h = ~c;
                if (!sb.ok_hit(h) || !is_prim(sb.hittables[h].kind)) { bad = true; return; }
                uint32_t ref = add_prim(h);
                fs.rank[prim_type(ref)][prim_index(ref)] = ref_rank[h];
                fs.leaf[prim_type(ref)][prim_index(ref)] = base + i;
                if (ref_sibling[h] >= 0) pending_siblings.push_back({ref, ref_sibling[h]});
            };
            if (n.left >= 0) walk(n.left); else leaf_child(n.left);
            if (n.right >= 0) walk(n.right); else if (!dup) leaf_child(n.right);
        };
        walk(b.root);
        for (auto& ps : pending_siblings) {
            if ((size_t)ps.second < prim_of.size() && prim_of[ps.second] != NO_PRIM)
                fs.sibling[prim_type(ps.first)][prim_index(ps.first)] = (int)prim_of[ps.second];
        }
        if (bad) { fail(SHIM_ERR_UNSUPPORTED_, "a BVH may only contain primitives and cubes"); return -1; }
        for (size_t i = 0; i < b.nodes.size(); ++i) {
            const HostBvhNode& n = b.nodes[i];
            DevNode dn;
            Box lb{}, rb{};
            int lref, rref;
            bool dup = n.left < 0 && n.right < 0 && n.left == n.right;
            auto child = [&](int c, Box& bx, int& ref) -> bool {
                if (c >= 0) { bx = b.nodes[c].box; ref = base + c; return true; }
                int h = ~c;
                if (!sb.ok_hit(h) || !is_prim(sb.hittables[h].kind)) return false;
                sb.bounding_box(h, b.t0, b.t1, bx);
                ref = (int)~add_prim(h);
                return true;
            };
            if (!child(n.left, lb, lref)) { fail(SHIM_ERR_UNSUPPORTED_, "a BVH may only contain primitives and cubes"); return -1; }
            if (dup) {
                rref = CHILD_NONE;
                for (int k = 0; k < 3; ++k) { rb.mn[k] = INFINITY; rb.mx[k] = -INFINITY; }
            } else if (!child(n.right, rb, rref)) { fail(SHIM_ERR_UNSUPPORTED_, "a BVH may only contain primitives and cubes"); return -1; }
            dn.a = f4{lb.mn[0], lb.mn[1], lb.mn[2], lb.mx[0]};
            dn.b = f4{lb.mx[1], lb.mx[2], rb.mn[0], rb.mn[1]};
            dn.c = f4{rb.mn[2], rb.mx[0], rb.mx[1], rb.mx[2]};
            dn.d = i4{lref, rref, n.parent >= 0 ? base + n.parent : -1, 0};
            fs.nodes[base + i] = dn;
        }
        bvh_base[hid] = base;
        bvh_root[hid] = b.root;
        bvh_nodes[hid] = (int)b.nodes.size();
        if (b.predictor) { bvh_pred[hid] = (int)fs.predictor_bvh.size(); fs.predictor_bvh.push_back(hid); }
        return base;
    }

    void emit(int hid, const Xf& xf) {
        const HostHittable& h = sb.hittables[hid];
        DevObject ob;
        memset(&ob, 0, sizeof ob);
        ob.predictor = -1;
        if (h.kind == H_BVH) {
            int base = add_bvh(hid);
            if (base < 0) return;
            ob.kind = OBJ_BVH; ob.ref = base + bvh_root[hid]; ob.n_nodes = bvh_nodes[hid];
            if (h.predictor) { ob.flags |= OBJ_PREDICTOR; ob.predictor = bvh_pred[hid]; }
        } else {
            ob.kind = OBJ_PRIM; ob.ref = (int)add_prim(hid);
        }
        if (xf.translate) { ob.flags |= OBJ_TRANSLATE; ob.dx = xf.d[0]; ob.dy = xf.d[1]; ob.dz = xf.d[2]; }
        if (xf.rotate) { ob.flags |= OBJ_ROTATE; }
        ob.sin_t = xf.s; ob.cos_t = xf.c;
        if (xf.medium >= 0) {
            const HostHittable& m = sb.hittables[xf.medium];
            ob.flags |= OBJ_MEDIUM; ob.handle = xf.medium; ob.neg_inv_density = m.neg_inv_density; ob.phase_mat = material_word(m.phase_mat);
        }
        fs.objects.push_back(ob);
    }

    void visit(int hid, Xf xf, int depth) {
        if (status) return;
        if (depth > 64) { fail(SHIM_ERR_UNSUPPORTED_, "hittable nesting too deep (cycle?)"); return; }
        const HostHittable& h = sb.hittables[hid];
        switch (h.kind) {
        case H_LIST:
            if (xf.medium >= 0) { fail(SHIM_ERR_UNSUPPORTED_, "ConstantMedium over a HittableList boundary is not supported"); return; }
            for (int it : h.items) visit(it, xf, depth + 1);
            return;
        case H_TRANSLATE:
            if (xf.translate || xf.rotate) { fail(SHIM_ERR_UNSUPPORTED_, "only Translate(RotateY(x)), Translate(x) and RotateY(x) instance chains are supported"); return; }
            xf.translate = true; xf.d[0] = h.p[0]; xf.d[1] = h.p[1]; xf.d[2] = h.p[2];
            visit(h.child, xf, depth + 1);
            return;
        case H_ROTATE_Y:
            if (xf.rotate) { fail(SHIM_ERR_UNSUPPORTED_, "only Translate(RotateY(x)), Translate(x) and RotateY(x) instance chains are supported"); return; }
            xf.rotate = true; xf.s = h.sin_t; xf.c = h.cos_t;
            visit(h.child, xf, depth + 1);
            return;
        case H_MEDIUM:
            if (xf.translate || xf.rotate || xf.medium >= 0) { fail(SHIM_ERR_UNSUPPORTED_, "a ConstantMedium must not sit under an instance transform or another medium"); return; }
            xf.medium = hid;
            visit(h.child, xf, depth + 1);
            return;
        default:
            emit(hid, xf);
            return;
        }
    }
};
}  // namespace

int SceneBuilder::flatten(FlatScene& fs) {
    fs = FlatScene();
    Flattener f{*this, fs};
    for (int w : world) f.visit(w, Flattener::Xf(), 0);
    if (f.status) return f.status;
    auto bits = [](int i) { float x; memcpy(&x, &i, 4); return x; };
    // textures
    for (const HostTexture& t : textures) {
        f4 a{bits(t.kind), bits(t.even), bits(t.odd), bits(0)}, b{t.color[0], t.color[1], t.color[2], t.scale};
        if (t.kind == TEX_MARBLE) { a.w = bits((int)fs.perlin.size()); marble_tables(t.seed, fs.perlin); }
        if (t.kind == TEX_IMAGE) {
            a.y = bits(t.w); a.z = bits(t.h); a.w = bits((int)fs.images.size());
            fs.images.insert(fs.images.end(), t.rgb.begin(), t.rgb.end());
        }
        fs.textures.push_back(a); fs.textures.push_back(b);
    }
    // materials (b.w = 1 when the texture chain reads u,v: only then is the sphere uv evaluated)
    std::function<bool(int, int)> reads_uv = [&](int t, int depth) -> bool {
        if (t < 0 || depth > 16) return false;
        if (textures[t].kind == TEX_IMAGE) return true;
        if (textures[t].kind == TEX_CHECKER) return reads_uv(textures[t].even, depth + 1) || reads_uv(textures[t].odd, depth + 1);
        return false;
    };
    for (const HostMaterial& m : materials) {
        fs.materials.push_back(f4{bits(m.kind), bits(m.tex), m.fuzz, m.ior});
        fs.materials.push_back(f4{m.albedo[0], m.albedo[1], m.albedo[2], reads_uv(m.tex, 0) ? 1.0f : 0.0f});
    }
    return 0;
}

SceneView FlatScene::view() const {
    SceneView v;
    memset(&v, 0, sizeof v);
    v.nodes = nodes.data(); v.snodes = snodes.empty() ? nullptr : snodes.data(); v.sph = sph.data(); v.sph_s = sph_s.data(); v.sph_mat = sph_mat.data();
    v.msph = msph.data(); v.rect = rect.data(); v.tri = tri.data(); v.cube = cube.data();
    v.objects = objects.data(); v.materials = materials.data(); v.textures = textures.data();
    v.images = images.data(); v.perlin = perlin.data();
    for (int i = 0; i < 5; ++i) { v.handle[i] = handle[i].data(); v.rank[i] = rank[i].data(); v.leaf[i] = leaf[i].data(); v.sibling[i] = sibling[i].data(); }
    v.n_objects = (int)objects.size(); v.n_nodes = (int)nodes.size();
    v.qnodes = qnodes.empty() ? nullptr : qnodes.data(); v.n_qnodes = (int)qnodes.size(); v.q_object = q_object;
    return v;
}
// DevNode -> SNode (shim_types.h): the planes of each axis in both orders, child node references as byte offsets
void FlatScene::build_signed_nodes() {
    snodes.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const DevNode& n = nodes[i];
        const float lmn[3] = {n.a.x, n.a.y, n.a.z}, lmx[3] = {n.a.w, n.b.x, n.b.y};
        const float rmn[3] = {n.b.z, n.b.w, n.c.x}, rmx[3] = {n.c.y, n.c.z, n.c.w};
        SNode& s = snodes[i];
        for (int k = 0; k < 3; ++k) {
            s.ax[2 * k] = f4{lmn[k], rmn[k], lmx[k], rmx[k]};
            s.ax[2 * k + 1] = f4{lmx[k], rmx[k], lmn[k], rmn[k]};
        }
        s.d = i4{n.d.x >= 0 ? n.d.x * (int)sizeof(SNode) : n.d.x, n.d.y >= 0 ? n.d.y * (int)sizeof(SNode) : n.d.y, n.d.z, 0};
    }
}

// DevNode -> QNode (shim_types.h) for the tree of object `oi`, renumbered breadth-first.  Everything is rounded
// outwards in exact (double) arithmetic against the very float the device decodes, plus one grid step of margin for
// the device's own f32 rounding of the plane distances.
namespace {
uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
float bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
double pow2(int s) { const uint64_t u = (uint64_t)(s + 1023) << 52; double d; memcpy(&d, &u, 8); return d; }   // -1022 <= s <= 1023
double floor_pos(double v) { return (double)(int64_t)v; }                                                    // 0 <= v < 2^62
double ceil_pos(double v) { const double f = (double)(int64_t)v; return f < v ? f + 1.0 : f; }
// grid of one axis for planes in [lo, hi]: origin bits (low byte = exponent byte of 2^15 * step) and the step
bool qgrid(float lo, float hi, uint32_t& o_bits, double& origin, double& step) {
    if (!std::isfinite(lo) || !std::isfinite(hi) || std::fabs(lo) > 4096.0f || std::fabs(hi) > 4096.0f || hi < lo) return false;
    const double ext = (double)hi - (double)lo;
    int s;
    if (ext > 0.0) {   // smallest s with 2^s >= ext / 252
        int e;
        const double m = std::frexp(ext / 252.0, &e);   // = m * 2^e, m in [0.5, 1)
        s = m == 0.5 ? e - 1 : e;
    } else {
        s = std::ilogb(std::fmax(std::fabs((double)lo), 1e-30)) - 20;
    }
    if (s < -126) s = -126;
    for (;; ++s) {
        const int E = s + 15 + 127;
        if (E < 1) continue;
        if (E > 254) return false;
        const double S = pow2(s);
        const double x = (double)lo - S;   // one step below the lowest plane
        // (float)x may round up by half an ulp, clearing the low bits moves a negative value up by < 512 ulps and the
        // exponent byte a positive one by < 256: one 512-ulp step down covers all of it (checked below)
        uint32_t b = (f32_bits((float)x) & ~0x1ffu) | (uint32_t)E;   // bit 8 stays clear: (word << 23) is then the step's float
        if ((double)bits_f32(b) > x) {      // one 512-ulp step towards -inf
            if (b & 0x80000000u) b += 512u;
            else if (b >= 1024u) b -= 512u;
            else b = 0x80000000u | (uint32_t)E;   // below the smallest positive grid value: a negative denormal
        }
        const double o = (double)bits_f32(b);
        if (!(o <= x) || !std::isfinite(o)) continue;
        if (ceil_pos(((double)hi - o) * pow2(-s)) + 1.0 > 255.0) continue;
        o_bits = b; origin = o; step = S;
        return true;
    }
}
}  // namespace
bool FlatScene::build_quantized_nodes(int oi) {
    qnodes.clear();
    q_object = -1;
    if (oi < 0 || oi >= (int)objects.size() || objects[oi].kind != OBJ_BVH) return false;
    std::vector<int> order{objects[oi].ref};   // QNode index -> DevNode index, breadth-first
    order.reserve((size_t)objects[oi].n_nodes);
    std::vector<int> qindex(nodes.size(), -1);
    qindex[objects[oi].ref] = 0;
    for (size_t h = 0; h < order.size(); ++h) {
        if (h + 16 < order.size()) __builtin_prefetch(&nodes[order[h + 16]].d);   // breadth-first order is random in memory
        const DevNode& n = nodes[order[h]];
        for (int c : {n.d.x, n.d.y})
            if (c >= 0) { qindex[c] = (int)order.size(); order.push_back(c); }
    }
    std::vector<QNode> out(order.size());
    // the nodes are independent: large trees (the commit of a 267 k-triangle mesh is inside the end-to-end time) are
    // quantised by a few threads
    std::atomic<bool> ok{true};
    auto quantise = [&](size_t begin, size_t end) {
      for (size_t i = begin; i < end; ++i) {
        if (i + 16 < end) __builtin_prefetch(&nodes[order[i + 16]]);
        const DevNode& n = nodes[order[i]];
        const bool has_r = n.d.y != CHILD_NONE;
        const float lmn[3] = {n.a.x, n.a.y, n.a.z}, lmx[3] = {n.a.w, n.b.x, n.b.y};
        const float rmn[3] = {n.b.z, n.b.w, n.c.x}, rmx[3] = {n.c.y, n.c.z, n.c.w};
        QNode& q = out[i];
        for (int k = 0; k < 3; ++k) {
            const float lo = has_r ? std::fmin(lmn[k], rmn[k]) : lmn[k], hi = has_r ? std::fmax(lmx[k], rmx[k]) : lmx[k];
            double origin, S;
            if (!qgrid(lo, hi, q.o[k], origin, S)) { ok = false; return; }
            const double inv_S = 1.0 / S;   // a power of two: exact
            auto down = [&](float v) { return (uint32_t)(floor_pos(((double)v - origin) * inv_S) - 1.0); };   // v - origin >= S
            auto up = [&](float v) { return (uint32_t)(ceil_pos(((double)v - origin) * inv_S) + 1.0); };
            if (!(lmn[k] <= lmx[k]) || (has_r && !(rmn[k] <= rmx[k]))) { ok = false; return; }
            q.q[k] = down(lmn[k]) | (has_r ? down(rmn[k]) : 255u) << 8 | up(lmx[k]) << 16 | (has_r ? up(rmx[k]) : 0u) << 24;
        }
        q.left = n.d.x >= 0 ? qindex[n.d.x] : n.d.x;
        q.right = n.d.y >= 0 ? qindex[n.d.y] : n.d.y;
      }
    };
    unsigned n_threads = order.size() >= 32768 ? std::thread::hardware_concurrency() : 1;
    if (n_threads > 8) n_threads = 8;
    if (n_threads <= 1) {
        quantise(0, order.size());
    } else {
        std::vector<std::thread> pool;
        const size_t per = (order.size() + n_threads - 1) / n_threads;
        for (unsigned t = 0; t < n_threads; ++t) {
            const size_t b = t * per, e = b + per < order.size() ? b + per : order.size();
            if (b < e) pool.emplace_back(quantise, b, e);
        }
        for (std::thread& t : pool) t.join();
    }
    if (!ok) return false;
    qnodes.swap(out);
    q_object = oi;
    return true;
}

uint64_t FlatScene::bytes() const {
    uint64_t b = nodes.size() * sizeof(DevNode) + snodes.size() * sizeof(SNode) + qnodes.size() * sizeof(QNode) + sph.size() * 8 + sph_s.size() * 16 + sph_mat.size() * 4 +
                 (msph.size() + rect.size() + tri.size() + cube.size() + materials.size() + textures.size()) * 16 +
                 objects.size() * sizeof(DevObject) + images.size() + perlin.size();
    for (int i = 0; i < 5; ++i) b += handle[i].size() * 16;
    return b;
}

// ---------------------------------------------------------------------------- camera.rs:44-81
namespace {
struct H3 { float x, y, z; };
H3 sub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
H3 mul(float s, H3 a) { return {s * a.x, s * a.y, s * a.z}; }
H3 divs(H3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
float dotp(H3 a, H3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
H3 crossp(H3 a, H3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
H3 norm(H3 a) { float inv = 1.0f / std::sqrt(dotp(a, a)); return {a.x * inv, a.y * inv, a.z * inv}; }
f3 tof3(H3 a) { f3 r; r.x = a.x; r.y = a.y; r.z = a.z; return r; }
}  // namespace

void camera_new(const float from[3], const float at[3], const float vup[3], float vfov, float aspect, float aperture,
                float focus_dist, float t0, float t1, CameraPod& out) {
    float theta = vfov * (kPi / 180.0f);  // f32::to_radians
    float h = std::tan(theta / 2.0f);
    float viewport_height = 2.0f * h;
    float viewport_width = aspect * viewport_height;
    H3 lf{from[0], from[1], from[2]}, la{at[0], at[1], at[2]}, up{vup[0], vup[1], vup[2]};
    H3 w = norm(sub(lf, la));
    H3 u = norm(crossp(up, w));
    H3 v = crossp(w, u);
    H3 horizontal = mul(focus_dist * viewport_width, u);
    H3 vertical = mul(focus_dist * viewport_height, v);
    H3 llc = sub(sub(sub(lf, divs(horizontal, 2.0f)), divs(vertical, 2.0f)), mul(focus_dist, w));
    out.origin = tof3(lf); out.horizontal = tof3(horizontal); out.vertical = tof3(vertical); out.llc = tof3(llc);
    out.u = tof3(u); out.v = tof3(v);
    out.lens_radius = aperture / 2.0f; out.time0 = t0; out.time1 = t1;
}

// ---------------------------------------------------------------------------- renderer.rs:242-296
std::vector<TileRect> tile_layout(int W, int H, int tw, int th) {
    std::vector<TileRect> tiles;
    int n_h = W / tw, rem_h = W % tw, n_v = H / th, rem_v = H % th;
    tiles.reserve((size_t)(n_h + 1) * (n_v + 1));
    for (int ty = 0; ty < n_v; ++ty) {
        for (int tx = 0; tx < n_h; ++tx) tiles.push_back({tw, th, tx * tw, ty * th});
        if (rem_h > 0) tiles.push_back({rem_h, th, n_h * tw, ty * th});
    }
    if (rem_v > 0)
        for (int tx = 0; tx < n_h; ++tx) tiles.push_back({tw, rem_v, tx * tw, n_v * th});
    if (rem_h > 0 && rem_v > 0) tiles.push_back({rem_h, rem_v, n_h * tw, n_v * th});
    return tiles;
}
std::vector<uint32_t> tile_pixel_order(int W, int H, int tw, int th, int rank, int world) {
    std::vector<TileRect> tiles = tile_layout(W, H, tw, th);
    std::vector<uint32_t> order;
    order.reserve((size_t)W * H / (world > 1 ? world : 1) + 64);
    for (size_t ti = 0; ti < tiles.size(); ++ti) {
        if (world > 1 && (int)(ti % (size_t)world) != rank) continue;
        const TileRect& t = tiles[ti];
        for (int y = 0; y < t.height; ++y)
            for (int x = 0; x < t.width; ++x) order.push_back((uint32_t)((t.y0 + y) * W + (t.x0 + x)));
    }
    return order;
}

}  // namespace shim
