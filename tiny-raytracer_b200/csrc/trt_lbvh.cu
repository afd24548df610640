// trt_lbvh.cu — BVH construction on the GPU (TRT_BVH_LBVH_DEVICE).
//
// Replaces the host builders for big scenes: BVH::new / Node::new (hittable/bvh.rs:12-21,42-84) is O(n log^2 n) on one
// core and the host SAH build of 10^6 primitives takes ~1 s, as long as rendering them. Here:
//   1. k_lbvh_bounds   bounds of the primitive-box centroids (block reduction + ordered-int atomics)
//   2. k_lbvh_morton   30-bit Morton code of every centroid
//   3. cub::DeviceRadixSort::SortPairs (code, primitive)
//   4. k_lbvh_karras   Karras 2012: every interior node finds its range and split from the sorted codes, in parallel
//   5. k_lbvh_refit    bottom-up boxes: the second child to arrive at a node merges (atomic flag per node)
//   6. k_lbvh_emit     the 64-byte node records of trt_layout.h (both children's boxes + child refs in the parent);
//                      also the tree depth (needed: the traversal stack holds TRT_MAX_STACK entries)
//   7. k_lbvh_collapse4  the 4-wide form of every node
//   8. k_ref_*         the reference's own Node::new, level by level (reference leaf ranks + the reference tree for the replay
//                      of order-dependent rays); see below
// The leaf boxes are the primitives' own padded boxes computed on the host with the reference's arithmetic
// (hittable/aabb.rs:13-20), interior boxes are exact min/max unions, so the culling argument of DESIGN.md §4.2 holds
// unchanged; the tree is walked near-child-first like the SAH tree.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include "trt_layout.h"
#include "trt_kernels.h"

namespace {

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

struct LbvhBounds { int mn[3], mx[3]; unsigned int depth; unsigned int pad; };

__global__ void k_lbvh_init(LbvhBounds* b) {
    for (int i = 0; i < 3; ++i) { b->mn[i] = 0x7FFFFFFF; b->mx[i] = (int)0x80000000; }
    b->depth = 0; b->pad = 0;
}

__global__ void k_lbvh_bounds(const float4* __restrict__ boxes, uint32_t n, LbvhBounds* b) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[3] = {0.f, 0.f, 0.f};
    bool ok = i < n;
    if (ok) {
        float4 lo = boxes[2 * (size_t)i], hi = boxes[2 * (size_t)i + 1];
        c[0] = 0.5f * (lo.x + hi.x); c[1] = 0.5f * (lo.y + hi.y); c[2] = 0.5f * (lo.z + hi.z);
    }
    for (int a = 0; a < 3; ++a) {
        int mn = ok ? f2ord(c[a]) : 0x7FFFFFFF, mx = ok ? f2ord(c[a]) : (int)0x80000000;
        for (int off = 16; off; off >>= 1) { mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, off)); mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, off)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&b->mn[a], mn); atomicMax(&b->mx[a], mx); }
    }
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void k_lbvh_morton(const float4* __restrict__ boxes, uint32_t n, const LbvhBounds* b, uint32_t* keys, uint32_t* vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo = boxes[2 * (size_t)i], hi = boxes[2 * (size_t)i + 1];
    float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    uint32_t q[3];
    for (int a = 0; a < 3; ++a) {
        float mn = ord2f(b->mn[a]), mx = ord2f(b->mx[a]);
        float ext = mx - mn;
        float u = ext > 0.f ? (c[a] - mn) / ext : 0.f;
        u = fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
        q[a] = (uint32_t)u;
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = i;
}

// common-prefix length of sorted keys i and j (index breaks ties), -1 outside the array
__device__ __forceinline__ int lbvh_delta(const uint32_t* __restrict__ k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = k[i], b = k[j];
    return a != b ? __clz(a ^ b) : 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

// child encoding inside the builder: bit31 = leaf (index into the sorted order), else interior node index
__global__ void k_lbvh_karras(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* left, uint32_t* right,
                              uint32_t* parent_of_node, uint32_t* parent_of_leaf, unsigned int* flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int ni = (int)n - 1;
    if (i >= ni) return;
    flags[i] = 0;
    int nn = (int)n;
    int d = lbvh_delta(keys, nn, i, i + 1) - lbvh_delta(keys, nn, i, i - 1) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(keys, nn, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, nn, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(keys, nn, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(keys, nn, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, nn, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    uint32_t L = (lo == gamma) ? (0x80000000u | (uint32_t)gamma) : (uint32_t)gamma;
    uint32_t R = (hi == gamma + 1) ? (0x80000000u | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    left[i] = L; right[i] = R;
    if (L & 0x80000000u) parent_of_leaf[L & 0x7FFFFFFFu] = (uint32_t)i; else parent_of_node[L] = (uint32_t)i;
    if (R & 0x80000000u) parent_of_leaf[R & 0x7FFFFFFFu] = (uint32_t)i; else parent_of_node[R] = (uint32_t)i;
    if (i == 0) parent_of_node[0] = 0xFFFFFFFFu;
}

__global__ void k_lbvh_refit(const float4* __restrict__ boxes, const uint32_t* __restrict__ order, uint32_t n,
                             const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                             const uint32_t* __restrict__ parent_of_node, const uint32_t* __restrict__ parent_of_leaf,
                             unsigned int* flags, float4* node_min, float4* node_max, LbvhBounds* b) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t cur = parent_of_leaf[i];
    {   // tree depth in levels = longest leaf-to-root chain (the traversal stack holds TRT_MAX_STACK entries)
        unsigned int depth = 1;
        for (uint32_t a = cur; a != 0xFFFFFFFFu; a = parent_of_node[a]) ++depth;
        const unsigned am = __activemask();
        depth = __reduce_max_sync(am, depth);
        if ((threadIdx.x & 31u) == (unsigned)(__ffs(am) - 1)) atomicMax(&b->depth, depth);
    }
    while (cur != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(&flags[cur], 1u) == 0u) return;       // first child to arrive: the sibling finishes this node
        __threadfence();
        float4 mn, mx;
        uint32_t ch[2] = {left[cur], right[cur]};
        for (int c = 0; c < 2; ++c) {
            float4 a, z;
            if (ch[c] & 0x80000000u) { uint32_t p = order[ch[c] & 0x7FFFFFFFu]; a = boxes[2 * (size_t)p]; z = boxes[2 * (size_t)p + 1]; }
            else { a = __ldcg(&node_min[ch[c]]); z = __ldcg(&node_max[ch[c]]); }   // written by another SM: read through L2
            if (c == 0) { mn = a; mx = z; }
            else { mn.x = fminf(mn.x, a.x); mn.y = fminf(mn.y, a.y); mn.z = fminf(mn.z, a.z); mx.x = fmaxf(mx.x, z.x); mx.y = fmaxf(mx.y, z.y); mx.z = fmaxf(mx.z, z.z); }
        }
        node_min[cur] = mn; node_max[cur] = mx;
        cur = parent_of_node[cur];
    }
}

__global__ void k_lbvh_emit(const float4* __restrict__ boxes, const uint32_t* __restrict__ order, uint32_t n,
                            const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                            const float4* __restrict__ node_min, const float4* __restrict__ node_max, float4* nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    uint32_t ch[2] = {left[i], right[i]};
    float4 mn[2], mx[2]; uint32_t ref[2];
    for (int c = 0; c < 2; ++c) {
        if (ch[c] & 0x80000000u) {
            uint32_t p = order[ch[c] & 0x7FFFFFFFu];
            mn[c] = boxes[2 * (size_t)p]; mx[c] = boxes[2 * (size_t)p + 1];
            ref[c] = __float_as_uint(mn[c].w);                  // the primitive's leaf ref travels in bmin.w
        } else { mn[c] = node_min[ch[c]]; mx[c] = node_max[ch[c]]; ref[c] = ch[c]; }
    }
    float4* rec = nodes + 4 * (size_t)i;
    rec[0] = make_float4(mn[0].x, mn[1].x, mn[0].y, mn[1].y);       // (left, right) pairs per coordinate, trt_layout.h
    rec[1] = make_float4(mn[0].z, mn[1].z, mx[0].x, mx[1].x);
    rec[2] = make_float4(mx[0].y, mx[1].y, mx[0].z, mx[1].z);
    rec[3] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), 0.f, 0.f);
}

// 4-wide record of binary node i (trt_scene.cpp collapse4 on the device): its children are node i's grandchildren, a leaf
// child stays as it is. Every binary node gets one — only those an even number of levels below the root are ever reached —
// so no compaction or depth is needed and a wide child ref is simply the binary index of that grandchild.
__global__ void k_lbvh_collapse4(const float4* __restrict__ nodes, uint32_t n, float4* nodes4, int greedy) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    float mn[4][3], mx[4][3]; uint32_t ref[4]; int nk = 0;
    auto put = [&](int k, const float4* r, int side) {
        const float4 a = r[0], b = r[1], c = r[2];
        mn[k][0] = side ? a.y : a.x; mn[k][1] = side ? a.w : a.z; mn[k][2] = side ? b.y : b.x;
        mx[k][0] = side ? b.w : b.z; mx[k][1] = side ? c.y : c.x; mx[k][2] = side ? c.w : c.z;
        ref[k] = __float_as_uint(side ? r[3].y : r[3].x);
    };
    const float4* rec = nodes + 4 * (size_t)i;
    if (!greedy) {
        for (int side = 0; side < 2; ++side) {
            uint32_t c = __float_as_uint(side ? rec[3].y : rec[3].x);
            if (c & 0x80000000u) put(nk++, rec, side);
            else { const float4* cr = nodes + 4 * (size_t)c; put(nk++, cr, 0); put(nk++, cr, 1); }
        }
    } else {
        // twice: open the interior child with the largest surface area, in place (like trt_scene.cpp collapse4 for the SAH tree)
        put(0, rec, 0); put(1, rec, 1); nk = 2;
        for (int rep = 0; rep < 2; ++rep) {
            int best = -1; float ba = -1.f;
            for (int k = 0; k < nk; ++k) if (!(ref[k] & 0x80000000u)) {
                const float dx = mx[k][0] - mn[k][0], dy = mx[k][1] - mn[k][1], dz = mx[k][2] - mn[k][2];
                const float a = dx * dy + dy * dz + dz * dx;
                if (a > ba) { ba = a; best = k; }
            }
            if (best < 0) break;
            const float4* cr = nodes + 4 * (size_t)ref[best];
            for (int k = nk; k > best + 1; --k) { for (int a = 0; a < 3; ++a) { mn[k][a] = mn[k - 1][a]; mx[k][a] = mx[k - 1][a]; } ref[k] = ref[k - 1]; }
            put(best, cr, 0); put(best + 1, cr, 1); ++nk;
        }
    }
    const float inf = __int_as_float(0x7f800000);
    for (; nk < 4; ++nk) { for (int a = 0; a < 3; ++a) { mn[nk][a] = inf; mx[nk][a] = -inf; } ref[nk] = 0x7FFFFFFFu; }
    float4* out = nodes4 + 8 * (size_t)i;
    for (int h = 0; h < 2; ++h) {
        const int a = 2 * h, b = 2 * h + 1;
        out[3 * h + 0] = make_float4(mn[a][0], mn[b][0], mn[a][1], mn[b][1]);
        out[3 * h + 1] = make_float4(mn[a][2], mn[b][2], mx[a][0], mx[b][0]);
        out[3 * h + 2] = make_float4(mx[a][1], mx[b][1], mx[a][2], mx[b][2]);
    }
    out[6] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), __uint_as_float(ref[2]), __uint_as_float(ref[3]));
    out[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ the REFERENCE tree on the device: Node::new, hittable/bvh.rs:42-84
// The device-built scene still needs the reference's left-to-right leaf order (exact-t ties go to the lower rank, and the
// rare order-dependent rays of DESIGN.md §4.3 are replayed in that order). Node::new is recursive, but every level of it is
// data-parallel: all slices of one depth are disjoint ranges of one permutation. Per level:
//   k_ref_bbox   union of the padded boxes of every slice of >= 2 primitives (ordered-int atomics, one per warp when a
//                warp lies inside one slice); the box is also stored into the parent's node record (the child side)
//   k_ref_keys   slices of >= 3: key = (slice start, f32::total_cmp key of bbox.min[longest axis]); slices of 1 or 2 keep
//                their order (bvh.rs:58-60): key = (slice start, position)
//   cub radix sort of (key, primitive): stable, and the slice start in the high word keeps every slice in place
//   k_ref_split  mid = len/2 (bvh.rs:75); the node of slice [lo,hi) is number mid-1 (every boundary between two positions
//                is the split of exactly one slice, so the n-1 nodes get distinct numbers); child refs and the boxes of
//                single-primitive children are written now, boxes of bigger children by the next level's k_ref_bbox
// ceil(log2 n) levels; 10^6 primitives: 20 sorts of 52-bit keys.
struct RefSeg { uint32_t lo, hi, parent, side; };   // per position: its slice, the node that owns the slice as a child, which side

__global__ void k_ref_init(uint32_t n, uint32_t* order, RefSeg* seg) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[i] = i;                                  // World insertion order (hittable/world.rs:23-26)
    seg[i].lo = 0; seg[i].hi = n; seg[i].parent = 0xFFFFFFFFu; seg[i].side = 0;
}
__global__ void k_ref_bbox_init(uint32_t n, int* bb) {   // bb[6n]: min x,y,z then max x,y,z, indexed by slice start
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int a = 0; a < 3; ++a) { bb[(size_t)a * n + i] = 0x7FFFFFFF; bb[(size_t)(3 + a) * n + i] = (int)0x80000000; }
}
__global__ void k_ref_bbox(const float4* __restrict__ boxes, const uint32_t* __restrict__ order, const RefSeg* __restrict__ seg, uint32_t n, int* bb) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < n;
    RefSeg s = ok ? seg[i] : RefSeg{0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0};
    const bool act = ok && s.hi - s.lo >= 2;
    int v[6];
    if (act) {
        const uint32_t p = order[i];
        const float4 lo = boxes[2 * (size_t)p], hi = boxes[2 * (size_t)p + 1];
        v[0] = f2ord(lo.x); v[1] = f2ord(lo.y); v[2] = f2ord(lo.z); v[3] = f2ord(hi.x); v[4] = f2ord(hi.y); v[5] = f2ord(hi.z);
    } else { v[0] = v[1] = v[2] = 0x7FFFFFFF; v[3] = v[4] = v[5] = (int)0x80000000; }
    const uint32_t lo0 = __shfl_sync(0xFFFFFFFFu, s.lo, 0);
    if (__all_sync(0xFFFFFFFFu, act && s.lo == lo0)) {           // the whole warp inside one slice: one atomic per plane
        for (int a = 0; a < 3; ++a) { v[a] = __reduce_min_sync(0xFFFFFFFFu, v[a]); v[3 + a] = __reduce_max_sync(0xFFFFFFFFu, v[3 + a]); }
        if ((threadIdx.x & 31) == 0)
            for (int a = 0; a < 3; ++a) { atomicMin(&bb[(size_t)a * n + lo0], v[a]); atomicMax(&bb[(size_t)(3 + a) * n + lo0], v[3 + a]); }
    } else if (act) {
        for (int a = 0; a < 3; ++a) { atomicMin(&bb[(size_t)a * n + s.lo], v[a]); atomicMax(&bb[(size_t)(3 + a) * n + s.lo], v[3 + a]); }
    }
}
// node record (trt_layout.h) of node `id`, child `side`: plane k (min x,y,z, max x,y,z) is float 2k + side, the ref float 12 + side
__device__ __forceinline__ void ref_put_child(float4* nodes, uint32_t id, uint32_t side, const float b[6], uint32_t ref) {
    float* r = (float*)(nodes + 4 * (size_t)id);
    for (int k = 0; k < 6; ++k) r[2 * k + side] = b[k];
    r[12 + side] = __uint_as_float(ref);
}
__global__ void k_ref_keys(const float4* __restrict__ boxes, const uint32_t* __restrict__ order, const RefSeg* __restrict__ seg, uint32_t n,
                           const int* __restrict__ bb, float4* nodes, unsigned long long* keys) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RefSeg s = seg[i];
    const uint32_t len = s.hi - s.lo;
    uint32_t low = i - s.lo;
    if (len >= 2) {
        float b[6];
        for (int k = 0; k < 6; ++k) b[k] = ord2f(bb[(size_t)k * n + s.lo]);
        if (i == s.lo && s.parent != 0xFFFFFFFFu) {              // this slice's box belongs into its parent's record
            const uint32_t clen = len, cmid = s.lo + (clen == 2 ? 1 : clen / 2);
            ref_put_child(nodes, s.parent, s.side, b, cmid - 1);
        }
        if (len >= 3) {
            const float s0 = b[3] - b[0], s1 = b[4] - b[1], s2 = b[5] - b[2];
            const int axis = (s0 > s1) ? ((s0 > s2) ? 0 : 2) : ((s1 > s2) ? 1 : 2);        // hittable/aabb.rs:63-78
            const float4 lo = boxes[2 * (size_t)order[i]];
            const float m = axis == 0 ? lo.x : (axis == 1 ? lo.y : lo.z);
            low = (uint32_t)f2ord(m) ^ 0x80000000u;                                          // total_cmp order as unsigned
        }
    }
    keys[i] = ((unsigned long long)s.lo << 32) | low;
}
__global__ void k_ref_split(const float4* __restrict__ boxes, const uint32_t* __restrict__ order, RefSeg* seg, uint32_t n, float4* nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RefSeg s = seg[i];
    const uint32_t len = s.hi - s.lo;
    if (len < 2) return;
    const uint32_t mid = s.lo + (len == 2 ? 1 : len / 2), id = mid - 1;
    const uint32_t side = i >= mid ? 1u : 0u;
    const uint32_t clo = side ? mid : s.lo, chi = side ? s.hi : mid;
    if (chi - clo == 1) {                                        // single-primitive child: its own padded box and leaf ref
        const uint32_t p = order[i];
        const float4 lo = boxes[2 * (size_t)p], hi = boxes[2 * (size_t)p + 1];
        const float b[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
        ref_put_child(nodes, id, side, b, __float_as_uint(lo.w));
    }
    s.lo = clo; s.hi = chi; s.parent = id; s.side = side;
    seg[i] = s;
}
// rank of every primitive = its final position; stored where the traversal reads it (meta.y, trt_layout.h)
__global__ void k_ref_ranks(const float4* __restrict__ boxes, const uint32_t* __restrict__ order, uint32_t n, uint4* sphere_meta, uint4* quad_meta) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = order[i];
    const uint32_t ref = __float_as_uint(boxes[2 * (size_t)p].w);
    uint4* m = (ref & TRT_REF_QUAD) ? quad_meta + (ref & TRT_REF_INDEX_MASK) : sphere_meta + (ref & TRT_REF_INDEX_MASK);
    m->y = i;
}


// ------------------------------------------------------------------ compressed 8-wide nodes (round 2)
// The walk of a tree that lives in L2 is bounded by the L1 data pipe (one wavefront per 32-byte sector per lane) and by the
// number of dependent fetches per ray. A wide-8 node holds up to eight children — what is left after opening, six times, the
// interior child with the largest surface area — with their boxes QUANTISED to 8 bits per plane on a per-node grid
// (origin + q * 2^e per axis, min planes rounded down, max planes rounded up: the quantised box contains the exact one), in
// 96 bytes = three 256-bit loads:
//   [0] origin.xyz, (biased exponent bytes ex | ey << 8 | ez << 16)      [1] qlo_x[8], qlo_y[8]     [2] qlo_z[8], qhi_x[8]
//   [3] qhi_y[8], qhi_z[8]                                               [4] child refs 0..3        [5] child refs 4..7
// A child ref is a leaf ref (trt_layout.h) or the index of a wide-8 node; absent children are TRT_REF_INVALID. The quantised
// boxes only CULL (conservatively); a leaf's own exact box is recomputed / fetched and tested with the reference's arithmetic
// before its primitive is tested (trt_device.cuh trav8_step), so results stay bit-identical.
// Built level by level from the binary records (the device-built tree or the uploaded SAH tree): one thread per wide node,
// interior children get consecutive indices from an atomic counter, `bin_of_wide` maps a wide index to its binary node.
__global__ void k_cw8_level(const float4* __restrict__ nodes, uint32_t* bin_of_wide, uint32_t begin, uint32_t end,
                            unsigned int* counter, float4* nodes8) {
    const uint32_t w = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= end) return;
    float mn[8][3], mx[8][3]; uint32_t ref[8]; int nk = 0;
    auto put = [&](int k, const float4* r, int side) {
        const float4 a = r[0], b = r[1], c = r[2];
        mn[k][0] = side ? a.y : a.x; mn[k][1] = side ? a.w : a.z; mn[k][2] = side ? b.y : b.x;
        mx[k][0] = side ? b.w : b.z; mx[k][1] = side ? c.y : c.x; mx[k][2] = side ? c.w : c.z;
        ref[k] = __float_as_uint(side ? r[3].y : r[3].x);
    };
    const float4* rec = nodes + 4 * (size_t)bin_of_wide[w];
    put(0, rec, 0); nk = 1;
    if (__float_as_uint(rec[3].y) != 0x7FFFFFFFu) { put(1, rec, 1); nk = 2; }
    while (nk < 8) {
        int best = -1; float ba = -1.f;
        for (int k = 0; k < nk; ++k) if (!(ref[k] & 0x80000000u)) {
            const float dx = mx[k][0] - mn[k][0], dy = mx[k][1] - mn[k][1], dz = mx[k][2] - mn[k][2];
            const float a = dx * dy + dy * dz + dz * dx;
            if (a > ba || best < 0) { ba = a; best = k; }
        }
        if (best < 0) break;
        const float4* cr = nodes + 4 * (size_t)ref[best];
        const bool two = __float_as_uint(cr[3].y) != 0x7FFFFFFFu;
        if (two) { for (int k = nk; k > best + 1; --k) { for (int a = 0; a < 3; ++a) { mn[k][a] = mn[k - 1][a]; mx[k][a] = mx[k - 1][a]; } ref[k] = ref[k - 1]; } }
        put(best, cr, 0);
        if (two) { put(best + 1, cr, 1); ++nk; }
    }
    int ni = 0;
    for (int k = 0; k < nk; ++k) if (!(ref[k] & 0x80000000u)) ++ni;
    const uint32_t base = ni ? atomicAdd(counter, (unsigned)ni) : 0u;
    for (int k = 0, j = 0; k < nk; ++k) if (!(ref[k] & 0x80000000u)) { bin_of_wide[base + j] = ref[k]; ref[k] = base + j; ++j; }
    float org[3]; int ee[3]; uint32_t qlo[3][2] = {{0, 0}, {0, 0}, {0, 0}}, qhi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int a = 0; a < 3; ++a) {
        float lo = mn[0][a], hi = mx[0][a];
        for (int k = 1; k < nk; ++k) { lo = fminf(lo, mn[k][a]); hi = fmaxf(hi, mx[k][a]); }
        org[a] = lo;
        const double ext = (double)hi - (double)lo;
        int e = -100;
        if (ext > 0.0) { frexp(ext / 255.0, &e); }               // ext / 255 = m * 2^e, m in [0.5, 1): 255 * 2^e > ext
        if (!(e >= -100)) e = -100;
        if (e > 100) e = 100;                                     // (never with |coordinates| <= 2^30, the only scenes this walk is used for)
        ee[a] = e;
        const double inv_step = ldexp(1.0, -e), step = ldexp(1.0, e);
        for (int k = 0; k < nk; ++k) {
            double ql = floor(((double)mn[k][a] - (double)lo) * inv_step), qh = ceil(((double)mx[k][a] - (double)lo) * inv_step);
            ql = fmin(fmax(ql, 0.0), 255.0); qh = fmin(fmax(qh, 0.0), 255.0);
            while (ql > 0.0 && (double)lo + ql * step > (double)mn[k][a]) ql -= 1.0;       // containment, whatever the rounding above did
            while (qh < 255.0 && (double)lo + qh * step < (double)mx[k][a]) qh += 1.0;
            qlo[a][k >> 2] |= (uint32_t)ql << (8 * (k & 3));
            qhi[a][k >> 2] |= (uint32_t)qh << (8 * (k & 3));
        }
    }
    for (int k = nk; k < 8; ++k) ref[k] = 0x7FFFFFFFu;
    float4* out = nodes8 + 6 * (size_t)w;
    const uint32_t eb = (uint32_t)(ee[0] + 127) | ((uint32_t)(ee[1] + 127) << 8) | ((uint32_t)(ee[2] + 127) << 16);
    out[0] = make_float4(org[0], org[1], org[2], __uint_as_float(eb));
    out[1] = make_float4(__uint_as_float(qlo[0][0]), __uint_as_float(qlo[0][1]), __uint_as_float(qlo[1][0]), __uint_as_float(qlo[1][1]));
    out[2] = make_float4(__uint_as_float(qlo[2][0]), __uint_as_float(qlo[2][1]), __uint_as_float(qhi[0][0]), __uint_as_float(qhi[0][1]));
    out[3] = make_float4(__uint_as_float(qhi[1][0]), __uint_as_float(qhi[1][1]), __uint_as_float(qhi[2][0]), __uint_as_float(qhi[2][1]));
    out[4] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), __uint_as_float(ref[2]), __uint_as_float(ref[3]));
    out[5] = make_float4(__uint_as_float(ref[4]), __uint_as_float(ref[5]), __uint_as_float(ref[6]), __uint_as_float(ref[7]));
}
__global__ void k_cw8_init(uint32_t* bin_of_wide, unsigned int* counter, uint32_t root) { bin_of_wide[0] = root; *counter = 1u; }

} // namespace

namespace trtk {

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

size_t lbvh_tmp_bytes(uint32_t n) {
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 30);
    // keys x2, vals x2, left, right, parent_of_node, parent_of_leaf, flags: 9 arrays of n u32; node_min/max: 2 arrays of n float4
    return al(sizeof(LbvhBounds)) + 9 * al(4ull * n) + 2 * al(16ull * n) + al(sort_tmp);
}

// d_boxes: 2 float4 per primitive (bmin.xyz + leaf-ref bits, bmax.xyz). d_nodes: 4 float4 per interior node (n - 1 nodes).
// d_nodes4 (optional): 8 float4 per interior node, the 4-wide record of every binary node (k_lbvh_collapse4).
cudaError_t build_lbvh(const void* d_boxes, uint32_t n, void* d_nodes, void* d_nodes4, void* d_tmp, uint32_t* h_depth, float* h_ms, cudaStream_t st) {
    // (the wide records are written after the depth is known: collapse_lbvh4)
    if (n < 2) return cudaErrorInvalidValue;
    char* p = (char*)d_tmp;
    LbvhBounds* b = (LbvhBounds*)p; p += al(sizeof(LbvhBounds));
    uint32_t* arr[9];
    for (int i = 0; i < 9; ++i) { arr[i] = (uint32_t*)p; p += al(4ull * n); }
    float4* node_min = (float4*)p; p += al(16ull * n);
    float4* node_max = (float4*)p; p += al(16ull * n);
    void* sort_tmp = p;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 30);
    uint32_t *keys = arr[0], *keys_s = arr[1], *vals = arr[2], *order = arr[3], *left = arr[4], *right = arr[5], *pn = arr[6], *pl = arr[7];
    unsigned int* flags = arr[8];
    const float4* boxes = (const float4*)d_boxes;
    cudaEvent_t e0, e1; cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return e; }
    unsigned g = (n + 255) / 256;
    cudaEventRecord(e0, st);
    k_lbvh_init<<<1, 1, 0, st>>>(b);
    k_lbvh_bounds<<<g, 256, 0, st>>>(boxes, n, b);
    k_lbvh_morton<<<g, 256, 0, st>>>(boxes, n, b, keys, vals);
    e = cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys, keys_s, vals, order, (int)n, 0, 30, st);
    if (e == cudaSuccess) {
        k_lbvh_karras<<<g, 256, 0, st>>>(keys_s, n, left, right, pn, pl, flags);
        k_lbvh_refit<<<g, 256, 0, st>>>(boxes, order, n, left, right, pn, pl, flags, node_min, node_max, b);
        k_lbvh_emit<<<g, 256, 0, st>>>(boxes, order, n, left, right, node_min, node_max, (float4*)d_nodes);
        cudaEventRecord(e1, st);
        LbvhBounds hb;
        e = cudaMemcpyAsync(&hb, b, sizeof hb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) { *h_depth = hb.depth; float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *h_ms = ms; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return e;
}

// wide records of a binary node array; greedy = open the largest children (a wide level may then consume one binary level only)
cudaError_t collapse_lbvh4(const void* d_nodes, uint32_t n, void* d_nodes4, bool greedy, cudaStream_t st) {
    k_lbvh_collapse4<<<(n + 255) / 256, 256, 0, st>>>((const float4*)d_nodes, n, (float4*)d_nodes4, greedy ? 1 : 0);
    return cudaGetLastError();
}


// Compressed 8-wide nodes from a binary node array (n - 1 records of 64 bytes, root index `root`): d_nodes8 must hold
// 96 * (n - 1) bytes (there are never more wide nodes than binary ones), d_tmp cw8_tmp_bytes(n). Synchronises the stream.
size_t cw8_tmp_bytes(uint32_t n) { return al(4ull * n) + 256; }
cudaError_t build_cw8(const void* d_nodes, uint32_t n, uint32_t root, void* d_nodes8, void* d_tmp, uint32_t* h_n_nodes8, uint32_t* h_levels, cudaStream_t st) {
    if (n < 2) return cudaErrorInvalidValue;
    uint32_t* bin_of_wide = (uint32_t*)d_tmp;
    unsigned int* counter = (unsigned int*)((char*)d_tmp + al(4ull * n));
    k_cw8_init<<<1, 1, 0, st>>>(bin_of_wide, counter, root);
    uint32_t begin = 0, end = 1, levels = 0;
    cudaError_t e = cudaSuccess;
    while (begin < end) {
        k_cw8_level<<<(end - begin + 127) / 128, 128, 0, st>>>((const float4*)d_nodes, bin_of_wide, begin, end, counter, (float4*)d_nodes8);
        unsigned int total = 0;
        e = cudaMemcpyAsync(&total, counter, sizeof total, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (total > n - 1 || total < end) return cudaErrorUnknown;
        begin = end; end = total; ++levels;
    }
    *h_n_nodes8 = end; *h_levels = levels;
    return cudaSuccess;
}

size_t ref_order_tmp_bytes(uint32_t n) {
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 64);
    return 2 * al(8ull * n) + 2 * al(4ull * n) + al(16ull * n) + al(24ull * n) + al(64ull * n) + al(sort_tmp);
}

// Node::new on the device (see the k_ref_* kernels): writes every primitive's reference leaf rank into meta.y and the
// reference tree's wide nodes (128 bytes per node, n - 1 nodes, root = *h_root) into d_nodes_ref4.
cudaError_t build_reference_order(const void* d_boxes, uint32_t n, void* d_sphere_meta, void* d_quad_meta, void* d_nodes_ref4,
                                  void* d_tmp, uint32_t* h_root, float* h_ms, cudaStream_t st) {
    if (n < 2) return cudaErrorInvalidValue;
    char* p = (char*)d_tmp;
    unsigned long long* keys = (unsigned long long*)p; p += al(8ull * n);
    unsigned long long* keys_s = (unsigned long long*)p; p += al(8ull * n);
    uint32_t* order = (uint32_t*)p; p += al(4ull * n);
    uint32_t* order_s = (uint32_t*)p; p += al(4ull * n);
    RefSeg* seg = (RefSeg*)p; p += al(16ull * n);
    int* bb = (int*)p; p += al(24ull * n);
    float4* nodes = (float4*)p; p += al(64ull * n);
    void* sort_tmp = p;
    size_t sort_bytes = 0;
    int seg_bits = 1; while ((1ull << seg_bits) < n) ++seg_bits;
    const int end_bit = 32 + seg_bits;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, end_bit);
    const float4* boxes = (const float4*)d_boxes;
    cudaEvent_t e0, e1; cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return e; }
    const unsigned g = (n + 255) / 256;
    cudaEventRecord(e0, st);
    k_ref_init<<<g, 256, 0, st>>>(n, order, seg);
    int levels = 0; while ((1ull << levels) < n) ++levels;            // ceil(log2 n): the longest slice halves (rounded up) per level
    for (int l = 0; l <= levels; ++l) {                                // one extra pass stores the boxes of the last two-primitive slices
        k_ref_bbox_init<<<g, 256, 0, st>>>(n, bb);
        k_ref_bbox<<<g, 256, 0, st>>>(boxes, order, seg, n, bb);
        k_ref_keys<<<g, 256, 0, st>>>(boxes, order, seg, n, bb, nodes, keys);
        if (l == levels) break;
        e = cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys, keys_s, order, order_s, (int)n, 0, end_bit, st);
        if (e != cudaSuccess) break;
        uint32_t* t = order; order = order_s; order_s = t;
        k_ref_split<<<g, 256, 0, st>>>(boxes, order, seg, n, nodes);
    }
    if (e == cudaSuccess) {
        k_ref_ranks<<<g, 256, 0, st>>>(boxes, order, n, (uint4*)d_sphere_meta, (uint4*)d_quad_meta);
        k_lbvh_collapse4<<<g, 256, 0, st>>>(nodes, n, (float4*)d_nodes_ref4, 0);
        cudaEventRecord(e1, st);
        e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) { float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *h_ms = ms; *h_root = (n == 2 ? 1u : n / 2) - 1u; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return e;
}

} // namespace trtk
