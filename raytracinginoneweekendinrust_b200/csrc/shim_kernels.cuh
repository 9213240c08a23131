// shim_kernels.cuh — sm_100a wavefront kernels.
//
// One iteration of the wavefront (replaces the per-tile loop of renderer.rs:63-85 and the
// recursion of ray.rs:32-62):
//
//   wf_begin     1 thread: tops the current ray queue up with new camera samples
//   wf_generate  camera rays for those samples (renderer.rs:141-143, camera.rs:96-106)
//   wf_extend    closest hit for every queued ray (hittable.rs:100-118); misses add the
//                background on the spot; hits are binned into per-material queues with one
//                atomic per warp and material
//   wf_shade<M>  one kernel per material kind: rebuild the HitRecord, emit, scatter, and
//                append the continuing ray to the next queue (warp-ballot compaction)
//
// Ray queues are SoA float4 streams in HBM, double buffered; all counts live on the device,
// the host only polls a done flag every few iterations.
#pragma once
#include <cuda_runtime.h>
#include "shim_device.h"

namespace shim {

enum { CNT_NRAYS0 = 0, CNT_NRAYS1 = 1, CNT_MQ = 2 /* ..6 */, CNT_GEN_BASE = 8, CNT_GEN_COUNT = 9, CNT_DONE = 10, CNT_ITER = 11,
       CNT_U64_BASE = 12 /* u64 slots from here, as pairs */ };
enum { C64_NEXT_SAMPLE = 0, C64_GEN_FIRST = 1, C64_RAYS = 2, C64_NODES = 3, C64_PRIMS = 4, C64_COUNT = 5 };
enum { CNT_WORDS = CNT_U64_BASE + 2 * C64_COUNT };

struct WfParams {
    SceneView sv;
    CameraPod cam;
    // ray queues (double buffered)
    f4* ray_o[2];   // origin.xyz, time
    f4* ray_d[2];   // direction.xyz, bounce (int bits)
    f4* thr[2];     // throughput.rgb, pixel index (int bits)
    uint32_t* samp[2];
    f4* hit;        // t, obj | face << 16, prim_ref, material
    uint32_t* mq[MAT_KINDS];
    uint32_t* cnt;
    float* accum;   // W*H*3 radiance sums
    const uint32_t* pix_table;
    uint32_t npix;            // pixels rendered by this call (tile shard)
    uint64_t total_samples;   // npix * sample_count
    uint32_t pool;
    int width, height, max_depth, sample_begin;
    float bg[3];
    uint64_t seed;
    int has_media, count_nodes;
};

__device__ __forceinline__ unsigned long long* cnt64(uint32_t* cnt, int slot) {
    return reinterpret_cast<unsigned long long*>(cnt + CNT_U64_BASE) + slot;
}

// position for this lane in a queue, one atomic per warp; all 32 lanes must call it
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool pred) {
    unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return 0;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

__global__ void wf_begin(WfParams p, int cur) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t* c = p.cnt;
    uint32_t n_cur = c[cur];
    unsigned long long next = *cnt64(c, C64_NEXT_SAMPLE);
    unsigned long long remaining = p.total_samples - next;
    unsigned long long room = (unsigned long long)(p.pool - n_cur);
    uint32_t n_new = (uint32_t)(remaining < room ? remaining : room);
    c[CNT_GEN_BASE] = n_cur;
    c[CNT_GEN_COUNT] = n_new;
    *cnt64(c, C64_GEN_FIRST) = next;
    *cnt64(c, C64_NEXT_SAMPLE) = next + n_new;
    c[cur] = n_cur + n_new;
    c[1 - cur] = 0;
#pragma unroll
    for (int k = 0; k < MAT_KINDS; ++k) c[CNT_MQ + k] = 0;
    *cnt64(c, C64_RAYS) += (unsigned long long)(n_cur + n_new);
    c[CNT_DONE] = (n_cur + n_new == 0) ? 1u : 0u;
    c[CNT_ITER] += (n_cur + n_new == 0) ? 0u : 1u;
}

__global__ void __launch_bounds__(256) wf_generate(WfParams p, int cur) {
    const uint32_t n = p.cnt[CNT_GEN_COUNT];
    const uint32_t base = p.cnt[CNT_GEN_BASE];
    const unsigned long long first = *cnt64(p.cnt, C64_GEN_FIRST);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        unsigned long long g = first + j;
        uint32_t s = (uint32_t)(g / p.npix);
        uint32_t pi = (uint32_t)(g - (unsigned long long)s * p.npix);
        uint32_t pixel = p.pix_table[pi];
        int x = (int)(pixel % (uint32_t)p.width), y = (int)(pixel / (uint32_t)p.width);
        uint32_t sample = (uint32_t)p.sample_begin + s;
        Rng rng;
        rng_init(rng, pixel, sample, p.seed);
        rng_key(rng, 0, STAGE_CAMERA);
        Ray r = camera_sample(p.cam, x, y, p.width, p.height, rng);
        uint32_t slot = base + j;
        f4 o; o.x = r.o.x; o.y = r.o.y; o.z = r.o.z; o.w = r.time;
        f4 d; d.x = r.d.x; d.y = r.d.y; d.z = r.d.z; d.w = i2f(0);
        f4 t; t.x = 1.0f; t.y = 1.0f; t.z = 1.0f; t.w = i2f((int)pixel);
        p.ray_o[cur][slot] = o;
        p.ray_d[cur][slot] = d;
        p.thr[cur][slot] = t;
        p.samp[cur][slot] = sample;
    }
}

__global__ void __launch_bounds__(256) wf_extend(WfParams p, int cur) {
    const uint32_t n = p.cnt[cur];
    const uint32_t n_round = (n + 31u) & ~31u;
    uint32_t nodes = 0, prims = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        const bool valid = i < n;
        int kind = -1;
        if (valid) {
            f4 o = p.ray_o[cur][i], d = p.ray_d[cur][i];
            Ray r; r.o = mk3(o.x, o.y, o.z); r.d = mk3(d.x, d.y, d.z); r.time = o.w;
            Rng rng;
            if (p.has_media) {
                rng_init(rng, (uint32_t)f2i(p.thr[cur][i].w), p.samp[cur][i], p.seed);
                rng_key(rng, (uint32_t)f2i(d.w), STAGE_INTERSECT);
            } else {
                rng_init(rng, 0, 0, 0);
            }
            TraceCounters tc; tc.nodes = 0; tc.prims = 0;
            Hit h = closest_hit(p.sv, r, 0.001f, SHIM_INF, rng, p.count_nodes ? &tc : nullptr);
            nodes += tc.nodes; prims += tc.prims;
            if (h.obj < 0) {  // ray.rs:60: miss returns the background
                if (p.bg[0] != 0.0f || p.bg[1] != 0.0f || p.bg[2] != 0.0f) {
                    f4 t = p.thr[cur][i];
                    float* a = p.accum + 3 * (size_t)(uint32_t)f2i(t.w);
                    atomicAdd(a + 0, t.x * p.bg[0]);
                    atomicAdd(a + 1, t.y * p.bg[1]);
                    atomicAdd(a + 2, t.z * p.bg[2]);
                }
            } else {
                int mat = hit_material(p.sv, h);
                kind = mat_kind(p.sv, mat);
                f4 hv; hv.x = h.t; hv.y = i2f(h.obj | (h.face << 16)); hv.z = i2f((int)h.prim); hv.w = i2f(mat);
                p.hit[i] = hv;
            }
        }
#pragma unroll
        for (int k = 0; k < MAT_KINDS; ++k) {
            uint32_t pos = warp_append(p.cnt + CNT_MQ + k, kind == k);
            if (kind == k) p.mq[k][pos] = i;
        }
    }
    if (p.count_nodes) {
        atomicAdd(cnt64(p.cnt, C64_NODES), (unsigned long long)nodes);
        atomicAdd(cnt64(p.cnt, C64_PRIMS), (unsigned long long)prims);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) wf_shade(WfParams p, int cur) {
    const int nxt = 1 - cur;
    const uint32_t n = p.cnt[CNT_MQ + KIND];
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x) {
        bool cont = false;
        Ray out; out.o = mk3(0, 0, 0); out.d = mk3(0, 0, 0); out.time = 0;
        f3 thr = mk3(0, 0, 0);
        int bounce = 0, pixel = 0; uint32_t sample = 0;
        if (j < n) {
            uint32_t i = p.mq[KIND][j];
            f4 o = p.ray_o[cur][i], d = p.ray_d[cur][i], t = p.thr[cur][i], hv = p.hit[i];
            sample = p.samp[cur][i];
            Ray r; r.o = mk3(o.x, o.y, o.z); r.d = mk3(d.x, d.y, d.z); r.time = o.w;
            bounce = f2i(d.w); pixel = f2i(t.w);
            thr = mk3(t.x, t.y, t.z);
            Hit h; h.t = hv.x; h.obj = f2i(hv.y) & 0xffff; h.face = f2i(hv.y) >> 16; h.prim = (uint32_t)f2i(hv.z);
            int mat = f2i(hv.w);
            HitRec rec;
            reconstruct_hit(p.sv, r, h, mat_needs_uv(p.sv, mat), rec);
            if (KIND == MAT_DIFFUSE_LIGHT) {  // ray.rs:46-48, 57: emitted, no scatter
                f3 e = mat_emit(p.sv, mat, rec);
                float* a = p.accum + 3 * (size_t)(uint32_t)pixel;
                atomicAdd(a + 0, thr.x * e.x);
                atomicAdd(a + 1, thr.y * e.y);
                atomicAdd(a + 2, thr.z * e.z);
            } else {
                Rng rng;
                rng_init(rng, (uint32_t)pixel, sample, p.seed);
                rng_key(rng, (uint32_t)bounce, STAGE_SCATTER);
                f3 att;
                if (mat_scatter(p.sv, KIND, mat, r, rec, rng, att, out)) {
                    thr = thr * att;
                    cont = bounce + 1 < p.max_depth;  // ray.rs:39-42: depth exhausted -> black
                }
            }
        }
        if (KIND != MAT_DIFFUSE_LIGHT) {
            uint32_t pos = warp_append(p.cnt + nxt, cont);
            if (cont) {
                f4 o; o.x = out.o.x; o.y = out.o.y; o.z = out.o.z; o.w = out.time;
                f4 d; d.x = out.d.x; d.y = out.d.y; d.z = out.d.z; d.w = i2f(bounce + 1);
                f4 t; t.x = thr.x; t.y = thr.y; t.z = thr.z; t.w = i2f(pixel);
                p.ray_o[nxt][pos] = o;
                p.ray_d[nxt][pos] = d;
                p.thr[nxt][pos] = t;
                p.samp[nxt][pos] = sample;
            }
        }
    }
}

// out = accum / spp (renderer.rs:147) or the raw sums
__global__ void wf_finalize(const float* accum, float* out, size_t n, float spp, int raw) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = raw ? accum[i] : accum[i] / spp;
}

// gate 1: closest hit for a batch of rays
__global__ void __launch_bounds__(256) trace_closest_kernel(SceneView sv, const float* rays, long long n, float t_min, float t_max,
                                                            uint64_t seed, int* prim_id, float* t_out, unsigned long long* counters) {
    uint32_t nodes = 0, prims = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float* q = rays + i * 7;
        Ray r; r.o = mk3(q[0], q[1], q[2]); r.d = mk3(q[3], q[4], q[5]); r.time = q[6];
        Rng rng;
        rng_init(rng, (uint32_t)i, 0, seed);
        rng_key(rng, 0, STAGE_INTERSECT);
        TraceCounters tc; tc.nodes = 0; tc.prims = 0;
        Hit h = closest_hit(sv, r, t_min, t_max, rng, counters ? &tc : nullptr);
        nodes += tc.nodes; prims += tc.prims;
        prim_id[i] = hit_handle(sv, h);
        t_out[i] = h.obj < 0 ? SHIM_INF : h.t;
    }
    if (counters) {
        atomicAdd(counters + 1, (unsigned long long)nodes);
        atomicAdd(counters + 2, (unsigned long long)prims);
    }
}

}  // namespace shim
