// shim_types.h — plain-old-data shared by the host flattener and the sm_100a kernels.
//
// HBM layout of a committed scene (see DESIGN.md "Data layout"):
//   nodes      DevNode[]      64 B/node: both child boxes + child refs + parent + leaf rank
//   sph        double[4*n]    cx,cy,cz,r*r   (extend: f64 quadratic, geometry/sphere.rs:57-83)
//   sph_s      f4[n]          cx,cy,cz,r + material in a side array (shade)
//   msph       f4[3*n]        {c0,r} {c1,time0} {time1, mat, -, -}
//   rect       f4[2*n]        {a0,a1,b0,b1} {k, axis, mat, -}
//   tri        f4[3*n]        {p0,mat} {e1,-} {e2,-}
//   cube       f4[2*n]        {min,mat} {max,-}
//   objects    DevObject[]    the flattened top-level HittableList, in list order
//   materials  f4[2*n], textures f4[2*n], image / perlin byte blobs
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SHIM_HD __host__ __device__ __forceinline__
#else
#define SHIM_HD inline
#endif

namespace shim {

struct f3 { float x, y, z; };
struct alignas(16) f4 { float x, y, z, w; };
struct alignas(16) i4 { int x, y, z, w; };

SHIM_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
SHIM_HD float i2f(int i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

enum PrimType { PT_SPHERE = 0, PT_MSPHERE = 1, PT_RECT = 2, PT_TRI = 3, PT_CUBE = 4, PT_NONE = 7 };
enum MatKind { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4, MAT_KINDS = 5 };
// Material queues are per shading CLASS: the five kinds plus one for Lambertians whose texture is expensive (marble:
// four f64 Perlin fBm evaluations, thousands of instructions).  In a queue shared with solid colours a warp ran that
// code for one or two lanes while the others waited (showcase: wf_shade at 9.7 of 32 lanes); in a queue of their
// own those hits fill whole warps.  The class travels in the kind field of a stored material reference; the trace
// pipeline (one-Bvh worlds) keeps the five kinds.
enum { MQ_SLOW_LAMBERTIAN = 5, MQ_CLASSES = 6 };
enum TexKind { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_MARBLE = 2, TEX_IMAGE = 3 };
enum { STAGE_CAMERA = 0, STAGE_INTERSECT = 1, STAGE_SCATTER = 2 };
// mirror shim_status in include/shimmer_b200.h
enum { SHIM_ERR_INVALID_ = -1, SHIM_ERR_UNSUPPORTED_ = -2, SHIM_ERR_CUDA_ = -3, SHIM_ERR_STATE_ = -4 };
enum { SHIM_MAX_BVH_HEIGHT = 38 };  // traversal stack is 40 entries

// primitive reference: type in the top 4 bits, index below
SHIM_HD uint32_t prim_ref(int type, uint32_t index) { return ((uint32_t)type << 28) | index; }
SHIM_HD int prim_type(uint32_t ref) { return (int)(ref >> 28); }
SHIM_HD uint32_t prim_index(uint32_t ref) { return ref & 0x0FFFFFFFu; }
static const int CHILD_NONE = (int)~(((uint32_t)PT_NONE << 28));  // empty child slot

// One BVH node = the reference's BvhNode (bvh.rs:229-236) with the same index, plus both
// CHILD boxes so one 64-byte fetch decides both descents.
struct alignas(16) DevNode {
    f4 a;  // l.min.x l.min.y l.min.z l.max.x
    f4 b;  // l.max.y l.max.z r.min.x r.min.y
    f4 c;  // r.min.z r.max.x r.max.y r.max.z
    i4 d;  // left ref, right ref (>=0 node, <0 ~prim_ref), parent, unused
};

// The same node for walks in shared memory (one-Bvh worlds): per axis the four child planes twice, once as
// {l.min, r.min, l.max, r.max} and once as {l.max, r.max, l.min, r.min}, so that ONE 16-byte load at offset 0 or 16 -
// chosen once per ray from the sign of its direction - returns {near_l, near_r, far_l, far_r} and the slab test needs
// no min/max to order the planes (12 fewer instructions per node; the ALU pipe is the busiest in the walk).
// Child references >= 0 are BYTE offsets of the child node (index * sizeof(SNode)), < 0 ~prim_ref as in DevNode.
struct alignas(16) SNode {
    f4 ax[6];  // [2 * axis + (direction negative ? 1 : 0)]
    i4 d;      // left ref, right ref, parent (node index), unused
};

// The node once more, quantised to 32 bytes, for the dense mesh walk (wf_bvh1_walk): one 32-byte fetch (LDG.256, or two
// LDS.128 for the top of the tree staged in shared memory) decides both descents, the node array is half as large (cache
// hit rates, a 4.9 k-triangle tree fits shared memory whole).  Per axis a grid origin o and a power-of-two step S; a
// child plane is o + q * S with an 8-bit q, rounded OUTWARDS by at least one step, so the decoded boxes contain the
// exact ones (a walk over larger boxes visits more candidates and returns the same hit, see bvh_closest).
//   o[k]   f32 bits of the origin; its low mantissa byte doubles as the IEEE exponent byte of 2^15 * S and bit 8 is
//          clear, so (o[k] << 23) is the float 2^15 * S (the origin is the float those 32 bits spell, exponent byte
//          included: the builder quantises against exactly that value)
//   q[k]   bytes {l.min, r.min, l.max, r.max}; an empty child slot is {.., 255, .., 0}: never hit
//   left/right   >= 0 index into the QNode array (breadth-first order, root = 0: the first K nodes are the top of the
//          tree), < 0 ~prim_ref, CHILD_NONE
struct alignas(32) QNode {
    uint32_t o[3];
    uint32_t q[3];
    int left, right;
};

enum ObjFlags { OBJ_TRANSLATE = 1, OBJ_ROTATE = 2, OBJ_MEDIUM = 4, OBJ_PREDICTOR = 8 };
enum ObjKind { OBJ_PRIM = 0, OBJ_BVH = 1 };

struct alignas(16) DevObject {
    int kind;      // OBJ_PRIM / OBJ_BVH
    int ref;       // prim_ref, or root node index
    int flags;
    int handle;    // user handle reported for medium hits
    float dx, dy, dz, sin_t;
    float cos_t, neg_inv_density;
    int phase_mat, predictor;
    int n_nodes, pad0, pad1, pad2;
};

struct HrppSlot;
struct SceneView {
    const DevNode* nodes;
    const SNode* snodes;    // the nodes again in the signed layout, or null (built for scenes that fit shared memory)
    const double* sph;      // 4 per sphere
    const f4* sph_s;
    const int* sph_mat;
    const f4* msph;         // 3 per moving sphere
    const f4* rect;         // 2 per rect
    const f4* tri;          // 3 per triangle
    const f4* cube;         // 2 per cube
    const DevObject* objects;
    const f4* materials;    // 2 per material
    const f4* textures;     // 2 per texture
    const uint8_t* images;
    const uint8_t* perlin;
    const int* handle[5];   // device prim index -> user handle, per PrimType
    const int* rank[5];     // device prim index -> left-to-right leaf rank inside its BVH (ties, bvh.rs:409-415)
    const int* leaf[5];     // device prim index -> node whose child it is (what HRPP stores, bvh.rs:382/398)
    const int* sibling[5];  // device prim index -> prim_ref of the other primitive of its recorded two-primitive leaf, or -1
    int n_objects;
    int n_nodes;
    // HRPP predictor tables (hrpp.rs:33-83), one per BVH that carries a predictor: open addressing,
    // hrpp_mask + 1 slots each
    struct HrppSlot* hrpp_slots;
    uint32_t hrpp_mask;
    int hrpp_log2;
    // quantised nodes of ONE Bvh object (q_object; the mesh of a one-Bvh-among-plain-objects world), or null
    const QNode* qnodes;
    int n_qnodes, q_object;
};
enum { HRPP_LEAVES = 4, HRPP_PROBES = 8 };
// One predictor slot = one 32-byte sector: the tagged 48-bit key and up to HRPP_LEAVES predicted leaf nodes, so a
// lookup that hits costs one memory transaction.  Empty = all ones (a table is cleared with one memset of 0xFF;
// a tag has bit 63 set and bits 48-62 clear, so it is never all ones).
struct alignas(32) HrppSlot {
    unsigned long long key;
    uint32_t leaf[HRPP_LEAVES];
    uint32_t pad[2];
};

struct CameraPod {  // camera.rs:6-27, derived on the host by Camera::new
    f3 origin, horizontal, vertical, llc, u, v;
    float lens_radius, time0, time1;
};

// closest-hit result, 16 B in the hit buffer
struct Hit {
    float t;
    int obj;        // top-level object index, -1 = miss
    uint32_t prim;  // prim_ref
    int face;       // cube side 0..5
};

}  // namespace shim
