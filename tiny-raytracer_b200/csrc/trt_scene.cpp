// trt_scene.cpp — World registry, BVH builders (reference-topology and binned SAH) and the HBM flattener.
// Host arithmetic that defines geometry (padded boxes, quad n/w/d) follows the reference operation for operation and
// is compiled without FMA contraction, because the kernels cull with exactly these numbers.
#include "trt_scene.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <thread>

namespace trt {

// Independent per-element loops over big scenes (10^6 primitives: the plain loops of a commit took 80 ms on one core)
template <class F> static void parallel_for(size_t n, F f) {
    const size_t min_chunk = 65536;
    size_t parts = n / min_chunk; if (parts > 8) parts = 8;
    unsigned hw = std::thread::hardware_concurrency(); if (hw && parts > hw) parts = hw;
    if (parts < 2) { f((size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t step = (n + parts - 1) / parts;
    for (size_t k = 1; k < parts; ++k) th.emplace_back([=] { f(k * step, std::min(n, (k + 1) * step)); });
    f((size_t)0, std::min(n, step));
    for (std::thread& t : th) t.join();
}

// ------------------------------------------------------------------ primitives
static inline void padded_box(const float a[3], const float b[3], float mn[3], float mx[3]) {
    // AABB::new, hittable/aabb.rs:13-20: PADDING_AMOUNT / 2.0 on each side
    const float pad = 0.0001f / 2.0f;
    for (int i = 0; i < 3; ++i) { mn[i] = fminf(a[i], b[i]) - pad; mx[i] = fmaxf(a[i], b[i]) + pad; }
}

uint32_t HostScene::add_material(uint32_t kind, const float rgb[3], float param) {
    HostMaterial m; m.kind = kind; m.rgb[0] = rgb[0]; m.rgb[1] = rgb[1]; m.rgb[2] = rgb[2]; m.param = param;
    if (kind == TRT_MAT_METAL) m.param = param < 0.f ? 0.f : (param > 1.f ? 1.f : param);   // material/metal.rs:13
    materials.push_back(m); committed = false;
    return (uint32_t)materials.size() - 1;
}

bool HostScene::add_sphere(const float c[3], float r, uint32_t mat, uint32_t* id, std::string* err) {
    if (mat >= materials.size()) { *err = "trt_add_sphere: unknown material id"; return false; }
    if (prims.size() >= TRT_REF_INDEX_MASK) { *err = "trt_add_sphere: too many primitives"; return false; }
    HostPrim p; std::memset(&p, 0, sizeof p);
    p.type = 0; p.mat = mat; p.r = r;
    float lo[3], hi[3];
    for (int i = 0; i < 3; ++i) { p.a[i] = c[i]; lo[i] = c[i] - r; hi[i] = c[i] + r; }        // hittable/sphere.rs:17-18
    padded_box(lo, hi, p.bmin, p.bmax);
    if (id) *id = (uint32_t)prims.size();
    prims.push_back(p); committed = false;
    return true;
}

bool HostScene::add_quad(const float corner[3], const float u[3], const float v[3], uint32_t mat, uint32_t* id, std::string* err) {
    if (mat >= materials.size()) { *err = "trt_add_quad: unknown material id"; return false; }
    if (prims.size() >= TRT_REF_INDEX_MASK) { *err = "trt_add_quad: too many primitives"; return false; }
    HostPrim p; std::memset(&p, 0, sizeof p);
    p.type = 1; p.mat = mat;
    float quv[3], qu[3], qv[3];
    for (int i = 0; i < 3; ++i) {
        p.a[i] = corner[i]; p.b[i] = u[i]; p.c[i] = v[i];
        quv[i] = (corner[i] + u[i]) + v[i]; qu[i] = corner[i] + u[i]; qv[i] = corner[i] + v[i];
    }
    // hittable/quad.rs:21-24: merge(new(corner, corner+u+v), new(corner+u, corner+v))
    float m0[3], x0[3], m1[3], x1[3];
    padded_box(corner, quv, m0, x0); padded_box(qu, qv, m1, x1);
    for (int i = 0; i < 3; ++i) { p.bmin[i] = fminf(m0[i], m1[i]); p.bmax[i] = fmaxf(x0[i], x1[i]); }
    // n = u x v; w = n / (n.n); d = n . corner   (hittable/quad.rs:25-27)
    p.n[0] = u[1] * v[2] - u[2] * v[1];
    p.n[1] = u[2] * v[0] - u[0] * v[2];
    p.n[2] = u[0] * v[1] - u[1] * v[0];
    float nn = p.n[0] * p.n[0] + p.n[1] * p.n[1] + p.n[2] * p.n[2];
    for (int i = 0; i < 3; ++i) p.w[i] = p.n[i] / nn;
    p.d = p.n[0] * corner[0] + p.n[1] * corner[1] + p.n[2] * corner[2];
    if (id) *id = (uint32_t)prims.size();
    prims.push_back(p); committed = false;
    return true;
}

// ------------------------------------------------------------------ generic binary tree produced by both builders
struct TNode {
    uint32_t left, right;    // indices into the TNode array; leaf when prim != ~0u
    uint32_t prim;
    float bmin[3], bmax[3];
};
static const uint32_t NOPRIM = 0xFFFFFFFFu;

static inline void box_union(TNode& n, const TNode& a, const TNode& b) {
    for (int i = 0; i < 3; ++i) { n.bmin[i] = fminf(a.bmin[i], b.bmin[i]); n.bmax[i] = fmaxf(a.bmax[i], b.bmax[i]); }
}
static inline int32_t total_order_key(float f) {   // f32::total_cmp (hittable/aabb.rs:80-82)
    int32_t k; std::memcpy(&k, &f, 4);
    k ^= (int32_t)(((uint32_t)(k >> 31)) >> 1);
    return k;
}

// ---- reference topology: Node::new, hittable/bvh.rs:42-84. The subtree over n primitives owns nodes [base, base+2n-1).
struct RefBuilder {
    const std::vector<HostPrim>& prims;
    std::vector<TNode>& nodes;
    std::vector<uint32_t>& order;   // working permutation, sorted in place slice by slice
    std::vector<int32_t> key[3];

    uint32_t build(size_t lo, size_t hi, uint32_t base, int par_depth) {
        size_t n = hi - lo;
        TNode& node = nodes[base];
        if (n == 1) {
            const HostPrim& p = prims[order[lo]];
            node.left = node.right = NOPRIM; node.prim = order[lo];
            for (int i = 0; i < 3; ++i) { node.bmin[i] = p.bmin[i]; node.bmax[i] = p.bmax[i]; }
            return base;
        }
        if (n > 2) {
            // union of the slice's boxes -> longest axis (hittable/aabb.rs:63-78), then stable sort on min[axis]
            float mn[3], mx[3];
            for (int i = 0; i < 3; ++i) { mn[i] = prims[order[lo]].bmin[i]; mx[i] = prims[order[lo]].bmax[i]; }
            for (size_t k = lo + 1; k < hi; ++k)
                for (int i = 0; i < 3; ++i) { mn[i] = fminf(mn[i], prims[order[k]].bmin[i]); mx[i] = fmaxf(mx[i], prims[order[k]].bmax[i]); }
            float s0 = mx[0] - mn[0], s1 = mx[1] - mn[1], s2 = mx[2] - mn[2];
            int axis = (s0 > s1) ? ((s0 > s2) ? 0 : 2) : ((s1 > s2) ? 1 : 2);
            const std::vector<int32_t>& kk = key[axis];
            std::stable_sort(order.begin() + lo, order.begin() + hi, [&kk](uint32_t a, uint32_t b) { return kk[a] < kk[b]; });
        }
        size_t nl = (n == 2) ? 1 : n / 2;                       // bvh.rs:58-60 (n == 2: unsorted), :75 (mid = len/2)
        uint32_t lbase = base + 1, rbase = base + (uint32_t)(2 * nl);
        uint32_t l, r;
        if (par_depth > 0 && n > 16384) {
            auto fut = std::async(std::launch::async, [&, lo, nl, lbase, par_depth]() { return build(lo, lo + nl, lbase, par_depth - 1); });
            r = build(lo + nl, hi, rbase, par_depth - 1);
            l = fut.get();
        } else {
            l = build(lo, lo + nl, lbase, 0);
            r = build(lo + nl, hi, rbase, 0);
        }
        node.left = l; node.right = r; node.prim = NOPRIM;
        box_union(node, nodes[l], nodes[r]);
        return base;
    }
};

// ---- binned SAH, one primitive per leaf. Same node numbering scheme.
// min/max without libm calls (-fno-fast-math keeps fminf/fmaxf out of line; the builder's inputs are never NaN)
static inline float fmin_(float a, float b) { return b < a ? b : a; }
static inline float fmax_(float a, float b) { return b > a ? b : a; }
struct SahBuilder {
    const std::vector<HostPrim>& prims;
    std::vector<TNode>& nodes;
    std::vector<uint32_t>& order;
    static constexpr int NB = 16;

    static float half_area(const float mn[3], const float mx[3]) {
        float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
    // The builder partitions an array of (box, id) records in place, so that a slice is contiguous in memory at every depth
    // (an index array into the primitives costs a cache miss per visit on big scenes: 10^6 spheres 900 -> 330 ms).
    struct Item { float b[6]; uint32_t id; };
    std::vector<Item> work;
    void prepare() {
        work.resize(prims.size());
        for (size_t i = 0; i < prims.size(); ++i) { for (int a = 0; a < 3; ++a) { work[i].b[a] = prims[i].bmin[a]; work[i].b[3 + a] = prims[i].bmax[a]; } work[i].id = (uint32_t)i; }
    }
    static float centroid(const Item& it, int ax) { return 0.5f * (it.b[ax] + it.b[3 + ax]); }

    uint32_t build(size_t lo, size_t hi, uint32_t base, int depth, int par_depth) {
        size_t n = hi - lo;
        TNode& node = nodes[base];
        if (n == 1) {
            const HostPrim& p = prims[work[lo].id];
            node.left = node.right = NOPRIM; node.prim = work[lo].id;
            for (int i = 0; i < 3; ++i) { node.bmin[i] = p.bmin[i]; node.bmax[i] = p.bmax[i]; }
            return base;
        }
        size_t mid = lo + n / 2;
        bool split_done = false;
        if (n > 2 && depth < TRT_MAX_STACK - 16) {
            float cmn[3], cmx[3];
            for (int a = 0; a < 3; ++a) { cmn[a] = std::numeric_limits<float>::infinity(); cmx[a] = -cmn[a]; }
            for (size_t k = lo; k < hi; ++k)
                for (int a = 0; a < 3; ++a) { float c = centroid(work[k], a); cmn[a] = fmin_(cmn[a], c); cmx[a] = fmax_(cmx[a], c); }
            float best_cost = std::numeric_limits<float>::infinity(); int best_axis = -1, best_bin = -1;
            // one pass over the slice bins all three axes at once
            float scale3[3]; bool use[3];
            for (int a = 0; a < 3; ++a) { float ext = cmx[a] - cmn[a]; use[a] = ext > 0.f; scale3[a] = use[a] ? (float)NB / ext : 0.f; }
            uint32_t cnt[3][NB]; float bmn[3][NB][3], bmx[3][NB][3];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < NB; ++b) { cnt[a][b] = 0; for (int i = 0; i < 3; ++i) { bmn[a][b][i] = std::numeric_limits<float>::infinity(); bmx[a][b][i] = -bmn[a][b][i]; } }
            for (size_t k = lo; k < hi; ++k) {
                const float* bx = work[k].b;
                for (int a = 0; a < 3; ++a) {
                    if (!use[a]) continue;
                    int b = (int)((0.5f * (bx[a] + bx[3 + a]) - cmn[a]) * scale3[a]); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
                    cnt[a][b]++;
                    for (int i = 0; i < 3; ++i) { bmn[a][b][i] = fmin_(bmn[a][b][i], bx[i]); bmx[a][b][i] = fmax_(bmx[a][b][i], bx[3 + i]); }
                }
            }
            for (int a = 0; a < 3; ++a) {
                if (!use[a]) continue;
                float la[NB - 1]; uint32_t lc[NB - 1];
                float amn[3], amx[3]; uint32_t c = 0;
                for (int i = 0; i < 3; ++i) { amn[i] = std::numeric_limits<float>::infinity(); amx[i] = -amn[i]; }
                for (int b = 0; b < NB - 1; ++b) {
                    c += cnt[a][b]; for (int i = 0; i < 3; ++i) { amn[i] = fmin_(amn[i], bmn[a][b][i]); amx[i] = fmax_(amx[i], bmx[a][b][i]); }
                    lc[b] = c; la[b] = c ? half_area(amn, amx) : 0.f;
                }
                c = 0; for (int i = 0; i < 3; ++i) { amn[i] = std::numeric_limits<float>::infinity(); amx[i] = -amn[i]; }
                for (int b = NB - 1; b >= 1; --b) {
                    c += cnt[a][b]; for (int i = 0; i < 3; ++i) { amn[i] = fmin_(amn[i], bmn[a][b][i]); amx[i] = fmax_(amx[i], bmx[a][b][i]); }
                    if (c == 0 || lc[b - 1] == 0) continue;
                    float cost = la[b - 1] * (float)lc[b - 1] + half_area(amn, amx) * (float)c;
                    if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b - 1; }
                }
            }
            if (best_axis >= 0) {
                float ext = cmx[best_axis] - cmn[best_axis], scale = (float)NB / ext, c0 = cmn[best_axis];
                int ax = best_axis, bb = best_bin;
                auto it = std::partition(work.begin() + lo, work.begin() + hi, [&](const Item& p) {
                    int b = (int)((centroid(p, ax) - c0) * scale); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
                    return b <= bb;
                });
                size_t m = (size_t)(it - work.begin());
                if (m > lo && m < hi) { mid = m; split_done = true; }
            }
        }
        if (!split_done && n > 2) {
            // degenerate centroids or depth cap: median on the widest box axis
            float mn[3], mx[3];
            for (int i = 0; i < 3; ++i) { mn[i] = std::numeric_limits<float>::infinity(); mx[i] = -mn[i]; }
            for (size_t k = lo; k < hi; ++k) for (int i = 0; i < 3; ++i) { mn[i] = fmin_(mn[i], work[k].b[i]); mx[i] = fmax_(mx[i], work[k].b[3 + i]); }
            int ax = 0; if (mx[1] - mn[1] > mx[ax] - mn[ax]) ax = 1; if (mx[2] - mn[2] > mx[ax] - mn[ax]) ax = 2;
            std::nth_element(work.begin() + lo, work.begin() + mid, work.begin() + hi, [&](const Item& a, const Item& b) { return centroid(a, ax) < centroid(b, ax); });
        }
        size_t nl = mid - lo;
        uint32_t lbase = base + 1, rbase = base + (uint32_t)(2 * nl);
        uint32_t l, r;
        if (par_depth > 0 && n > 16384) {
            auto fut = std::async(std::launch::async, [&, lo, mid, lbase, depth, par_depth]() { return build(lo, mid, lbase, depth + 1, par_depth - 1); });
            r = build(mid, hi, rbase, depth + 1, par_depth - 1);
            l = fut.get();
        } else {
            l = build(lo, mid, lbase, depth + 1, 0);
            r = build(mid, hi, rbase, depth + 1, 0);
        }
        node.left = l; node.right = r; node.prim = NOPRIM;
        box_union(node, nodes[l], nodes[r]);
        return base;
    }
};

// ------------------------------------------------------------------ flatten
static inline float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static uint32_t leaf_ref(const HostPrim& p) {
    return TRT_REF_LEAF | (p.type ? TRT_REF_QUAD : 0u) | p.type_index;
}

static void flatten(const std::vector<TNode>& t, const std::vector<HostPrim>& prims, FlatScene& out) {
    // interior nodes in DFS preorder (parent, whole left subtree, whole right subtree): a child is near its parent in HBM
    size_t n_leaves = prims.size();
    out.nodes.clear();
    out.root = TRT_REF_INVALID; out.n_nodes = 0; out.max_depth = 0; out.sah_cost = 0.f;
    if (n_leaves == 0) return;
    auto put_box = [](trt_f4* rec, const TNode* l, const TNode* r) {
        const float inf = std::numeric_limits<float>::infinity();
        float lmn[3] = {inf, inf, inf}, lmx[3] = {-inf, -inf, -inf}, rmn[3] = {inf, inf, inf}, rmx[3] = {-inf, -inf, -inf};
        if (l) for (int i = 0; i < 3; ++i) { lmn[i] = l->bmin[i]; lmx[i] = l->bmax[i]; }
        if (r) for (int i = 0; i < 3; ++i) { rmn[i] = r->bmin[i]; rmx[i] = r->bmax[i]; }
        rec[0] = {lmn[0], rmn[0], lmn[1], rmn[1]};       // (left, right) pairs per coordinate, trt_layout.h
        rec[1] = {lmn[2], rmn[2], lmx[0], rmx[0]};
        rec[2] = {lmx[1], rmx[1], lmx[2], rmx[2]};
    };
    if (n_leaves == 1) {   // a lone primitive: pseudo-node whose right child is absent
        out.nodes.resize(4);
        put_box(out.nodes.data(), &t[0], nullptr);
        out.nodes[3] = {as_float(leaf_ref(prims[t[0].prim])), as_float(TRT_REF_INVALID), 0.f, 0.f};
        out.root = 0; out.n_nodes = 1; out.max_depth = 1;
        return;
    }
    out.nodes.resize(4 * (n_leaves - 1));
    struct Item { uint32_t tnode, depth; };
    std::vector<uint32_t> flat_index(t.size(), 0xFFFFFFFFu);
    // pass 1: assign preorder indices to interior nodes
    {
        uint32_t next = 0; std::vector<Item> st; st.push_back({0u, 1u});
        while (!st.empty()) {
            Item it = st.back(); st.pop_back();
            const TNode& n = t[it.tnode];
            if (n.prim != NOPRIM) { if (it.depth > out.max_depth) out.max_depth = it.depth; continue; }
            flat_index[it.tnode] = next++;
            st.push_back({n.right, it.depth + 1}); st.push_back({n.left, it.depth + 1});
        }
        out.n_nodes = next;
    }
    auto ha = [](const TNode& n) { float dx = n.bmax[0] - n.bmin[0], dy = n.bmax[1] - n.bmin[1], dz = n.bmax[2] - n.bmin[2]; return dx * dy + dy * dz + dz * dx; };
    const double root_area = ha(t[0]) > 0.f ? ha(t[0]) : 1.0;
    std::vector<uint32_t> refs(prims.size());                 // 4 bytes per primitive instead of a HostPrim per lookup
    for (size_t i = 0; i < prims.size(); ++i) refs[i] = leaf_ref(prims[i]);
    // pass 2: the records; independent per node, so big trees are written by a few threads
    auto emit = [&](size_t i0, size_t i1) -> double {
        double c = 0.0;
        for (size_t i = i0; i < i1; ++i) {
            const TNode& n = t[i];
            c += (double)ha(n) / root_area;
            if (n.prim != NOPRIM) continue;
            const TNode& l = t[n.left]; const TNode& r = t[n.right];
            trt_f4* rec = &out.nodes[4 * (size_t)flat_index[i]];
            put_box(rec, &l, &r);
            uint32_t lr = (l.prim != NOPRIM) ? refs[l.prim] : flat_index[n.left];
            uint32_t rr = (r.prim != NOPRIM) ? refs[r.prim] : flat_index[n.right];
            rec[3] = {as_float(lr), as_float(rr), 0.f, 0.f};
        }
        return c;
    };
    double cost = 0.0;
    const size_t parts = t.size() > 200000 ? 8 : 1, step = (t.size() + parts - 1) / parts;
    std::vector<std::future<double>> fut;
    for (size_t k = 1; k < parts; ++k) fut.push_back(std::async(std::launch::async, emit, k * step, std::min(t.size(), (k + 1) * step)));
    std::vector<double> part_cost(parts, 0.0);
    part_cost[0] = emit(0, std::min(t.size(), step));
    for (size_t k = 1; k < parts; ++k) part_cost[k] = fut[k - 1].get();
    for (double c : part_cost) cost += c;                     // fixed summation order: the reported cost does not depend on scheduling
    out.root = 0; out.sah_cost = (float)cost;
}

// 4-wide nodes: a wide node's children are the grandchildren of a binary node (a child that is a leaf stays as it is) or,
// for the SAH tree (greedy), what is left after twice opening the interior child with the largest surface area. Either
// way every box is still one of the binary tree's boxes — in particular a leaf's box is the primitive's own padded box —
// left-to-right order is kept, and the traversal does about half as many dependent node fetches per ray. Record = eight
// float4: [0..2] children 0,1 and [3..5] children 2,3 as (a, b) pairs per coordinate like the binary node (trt_layout.h),
// [6] the four child refs, [7] unused. Absent children: TRT_REF_INVALID with an empty box. DFS preorder.
static void collapse4(const std::vector<TNode>& t, const std::vector<HostPrim>& prims, FlatScene& out, bool greedy = false) {
    out.nodes4.clear(); out.n_nodes4 = 0; out.max_depth4 = 0;
    if (prims.size() < 2) return;
    struct Item { uint32_t tnode, index, depth; };
    std::vector<Item> st;
    uint32_t next = 1;
    st.push_back({0u, 0u, 1u});
    const float inf = std::numeric_limits<float>::infinity();
    while (!st.empty()) {
        Item it = st.back(); st.pop_back();
        if (it.depth > out.max_depth4) out.max_depth4 = it.depth;
        const TNode& n = t[it.tnode];
        uint32_t kids[4]; int nk = 0;
        if (!greedy) {
            for (uint32_t c : {n.left, n.right}) {
                if (t[c].prim != NOPRIM) kids[nk++] = c;
                else { kids[nk++] = t[c].left; kids[nk++] = t[c].right; }
            }
        } else {
            // open the interior child with the largest surface area, twice, in place (left-to-right order is kept)
            kids[nk++] = n.left; kids[nk++] = n.right;
            auto area = [&](uint32_t c) { const TNode& k = t[c]; float dx = k.bmax[0] - k.bmin[0], dy = k.bmax[1] - k.bmin[1], dz = k.bmax[2] - k.bmin[2]; return dx * dy + dy * dz + dz * dx; };
            for (int rep = 0; rep < 2; ++rep) {
                int best = -1; float ba = -1.f;
                for (int k = 0; k < nk; ++k) if (t[kids[k]].prim == NOPRIM) { float a = area(kids[k]); if (a > ba) { ba = a; best = k; } }
                if (best < 0) break;
                const uint32_t c = kids[best];
                for (int k = nk; k > best + 1; --k) kids[k] = kids[k - 1];
                kids[best] = t[c].left; kids[best + 1] = t[c].right; ++nk;
            }
        }
        if (out.nodes4.size() < 8 * (size_t)(it.index + 1)) out.nodes4.resize(8 * (size_t)(it.index + 1));
        float mn[4][3], mx[4][3]; uint32_t ref[4];
        uint32_t child_index[4] = {0, 0, 0, 0};
        for (int k = 0; k < 4; ++k) {
            if (k < nk) {
                const TNode& c = t[kids[k]];
                for (int a = 0; a < 3; ++a) { mn[k][a] = c.bmin[a]; mx[k][a] = c.bmax[a]; }
                if (c.prim != NOPRIM) ref[k] = leaf_ref(prims[c.prim]);
                else { child_index[k] = next++; ref[k] = child_index[k]; }
            } else { for (int a = 0; a < 3; ++a) { mn[k][a] = inf; mx[k][a] = -inf; } ref[k] = TRT_REF_INVALID; }
        }
        // children are pushed in reverse so that the first child's subtree follows its parent in memory
        for (int k = nk - 1; k >= 0; --k) if (t[kids[k]].prim == NOPRIM) st.push_back({kids[k], child_index[k], it.depth + 1});
        if (out.nodes4.size() < 8 * (size_t)next) out.nodes4.resize(8 * (size_t)next);
        trt_f4* rec = &out.nodes4[8 * (size_t)it.index];
        for (int h = 0; h < 2; ++h) {
            const int a = 2 * h, b = 2 * h + 1;
            rec[3 * h + 0] = {mn[a][0], mn[b][0], mn[a][1], mn[b][1]};
            rec[3 * h + 1] = {mn[a][2], mn[b][2], mx[a][0], mx[b][0]};
            rec[3 * h + 2] = {mx[a][1], mx[b][1], mx[a][2], mx[b][2]};
        }
        rec[6] = {as_float(ref[0]), as_float(ref[1]), as_float(ref[2]), as_float(ref[3])};
        rec[7] = {0.f, 0.f, 0.f, 0.f};
    }
    out.n_nodes4 = next;
    out.nodes4.resize(8 * (size_t)next);
}

bool HostScene::commit(int bvh_kind, std::string* err) {
    if (bvh_kind < 0 || bvh_kind > 4) { *err = "trt_scene_commit: unknown bvh_kind"; return false; }
    size_t n = prims.size();
    if (bvh_kind == 4 && n < 2) bvh_kind = 0;                    // nothing to build on the device
    // TRT_BVH_AUTO: up to TRT_FLAT_MAX primitives the flat list beats any tree on a SIMT machine (measured: Cornell 15.8 vs
    // 10.8 Grays/s, glass Cornell 9.8 vs 6.3) and reproduces the reference by construction; larger scenes need SAH.
    if (bvh_kind == 3) bvh_kind = (n >= 1 && n <= TRT_FLAT_MAX) ? 2 : (n == 0 ? 0 : (n >= TRT_AUTO_DEVICE_BUILD_MIN ? 4 : 1));
    if (bvh_kind == 2 && (n < 1 || n > TRT_FLAT_MAX)) { *err = "trt_scene_commit: TRT_BVH_FLAT takes 1..64 primitives"; return false; }
    // per-type arrays in insertion order (what World::get_geometries::<T>() returns, hittable/world.rs:47-55)
    uint32_t ns = 0, nq = 0;
    for (auto& p : prims) p.type_index = p.type ? nq++ : ns++;
    leaf_rank.assign(n, 0);
    flat.reset(bvh_kind);

    const bool timing = getenv("TRT_COMMIT_TIMING") != nullptr;      // stderr: where the commit time goes
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "trt_scene_commit: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    const bool device_build = bvh_kind == 4;
    // big SAH scenes: the reference tree (leaf ranks, replay tree) is left to the device as well (trt_lbvh.cu k_ref_*), which
    // does in milliseconds what takes the host half of its commit time
    const bool ref_on_device = bvh_kind == 1 && n >= TRT_AUTO_DEVICE_BUILD_MIN;
    flat.ref_on_device = ref_on_device;
    std::vector<TNode> ref_tree(n && !device_build && !ref_on_device ? 2 * n - 1 : 0), sah_tree;
    std::vector<uint32_t> order(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    if (device_build || ref_on_device) {
        // TRT_BVH_LBVH_DEVICE: no host tree at all. The node arrays AND the reference leaf ranks (meta.y, placeholder here)
        // are produced by trt_lbvh.cu during trt_scene_upload from these per-primitive boxes.
        flat.lbvh_boxes.resize(2 * n);
        parallel_for(n, [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) {
                const HostPrim& p = prims[i];
                leaf_rank[i] = (uint32_t)i;
                flat.lbvh_boxes[2 * i] = {p.bmin[0], p.bmin[1], p.bmin[2], as_float(leaf_ref(p))};
                flat.lbvh_boxes[2 * i + 1] = {p.bmax[0], p.bmax[1], p.bmax[2], 0.f};
            }
        });
        if (device_build) { flat.root = 0; flat.n_nodes = (uint32_t)(n - 1); flat.max_depth = 0; flat.sah_cost = 0.f; }
    } else if (n) {
        RefBuilder rb{prims, ref_tree, order, {}};
        for (int a = 0; a < 3; ++a) { rb.key[a].resize(n); for (size_t i = 0; i < n; ++i) rb.key[a][i] = total_order_key(prims[i].bmin[a]); }
        rb.build(0, n, 0, 3);
        // the in-place sorts leave `order` in left-to-right leaf order
        for (size_t i = 0; i < n; ++i) leaf_rank[order[i]] = (uint32_t)i;
    }
    lap("reference tree / boxes");
    const std::vector<TNode>* tree = &ref_tree;
    if (bvh_kind == 1 && n) {
        sah_tree.resize(2 * n - 1);
        for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
        SahBuilder sb{prims, sah_tree, order, {}};
        sb.prepare();
        sb.build(0, n, 0, 1, 4);
        tree = &sah_tree;
    }
    lap("SAH tree");
    if (!device_build) flatten(*tree, prims, flat);
    lap("flatten");
    flat.wide_on_device = ref_on_device && 3 * ((flat.max_depth + 1) / 2) + 1 <= TRT_MAX_STACK;   // big SAH scene: k_lbvh_collapse4 during upload
    if ((bvh_kind == 0 || bvh_kind == 1) && n && !flat.wide_on_device) {
        collapse4(*tree, prims, flat, bvh_kind == 1);     // SAH: open the largest children (final scene +3 %)
        if (3 * flat.max_depth4 + 1 > TRT_MAX_STACK) { flat.nodes4.clear(); flat.n_nodes4 = 0; }   // walk the binary tree instead
    }
    if (bvh_kind == 1 && n && !ref_on_device) {   // the reference-topology tree too: the ordered walk replays its rare order-dependent rays on it
        FlatScene tmp;
        collapse4(ref_tree, prims, tmp);
        if (3 * tmp.max_depth4 + 1 <= TRT_MAX_STACK) { flat.nodes_ref4.swap(tmp.nodes4); flat.n_nodes_ref4 = tmp.n_nodes4; }
    }
    lap("wide nodes");
    if (flat.max_depth > TRT_MAX_STACK) { *err = "trt_scene_commit: tree deeper than the traversal stack"; return false; }

    flat.spheres.resize(ns); flat.sphere_meta.resize(ns); flat.quads.resize(4 * (size_t)nq); flat.quad_nn.resize(nq); flat.quad_meta.resize(nq);
    parallel_for(n, [&](size_t i0, size_t i1) {       // every primitive writes its own slots of its type's arrays
        for (size_t i = i0; i < i1; ++i) {
            const HostPrim& p = prims[i];
            trt_u4 meta = {p.mat, leaf_rank[i], (uint32_t)i, 0u};
            if (p.type == 0) {
                flat.spheres[p.type_index] = {p.a[0], p.a[1], p.a[2], p.r};
                flat.sphere_meta[p.type_index] = meta;
            } else {
                trt_f4* q = &flat.quads[4 * (size_t)p.type_index];
                q[0] = {p.n[0], p.n[1], p.n[2], p.d};
                q[1] = {p.a[0], p.a[1], p.a[2], p.w[0]};
                q[2] = {p.b[0], p.b[1], p.b[2], p.w[1]};
                q[3] = {p.c[0], p.c[1], p.c[2], p.w[2]};
                flat.quad_meta[p.type_index] = meta;
                // outward_normal.normalized() of hittable/mod.rs:36-40 for this quad: v / sqrt(v.v), true divisions (vec3.rs:45-47)
                float len = std::sqrt(p.n[0] * p.n[0] + p.n[1] * p.n[1] + p.n[2] * p.n[2]);
                flat.quad_nn[p.type_index] = {p.n[0] / len, p.n[1] / len, p.n[2] / len, 0.f};
            }
        }
    });
    if (bvh_kind == 2) {
        // the list, in reference leaf order: geometry, then the primitive's own padded box and its child-ref bits
        flat.flat.assign(TRT_FLAT_F4 * n, trt_f4{0.f, 0.f, 0.f, 0.f});
        flat.flat_qmask[0] = flat.flat_qmask[1] = 0u;
        for (size_t i = 0; i < n; ++i) {
            const HostPrim& p = prims[i];
            const size_t r = leaf_rank[i];
            trt_f4* f = &flat.flat[TRT_FLAT_F4 * r];
            if (p.type == 0) { f[0] = {p.a[0], p.a[1], p.a[2], p.r}; f[1] = f[2] = f[3] = {0.f, 0.f, 0.f, 0.f}; }
            else { for (int k = 0; k < 4; ++k) f[k] = flat.quads[4 * (size_t)p.type_index + k]; flat.flat_qmask[r >> 5] |= 1u << (r & 31); }
            f[4] = {p.bmin[0], p.bmax[0], p.bmin[1], p.bmax[1]};
            f[5] = {p.bmin[2], p.bmax[2], as_float(leaf_ref(p)), 0.f};
        }
        // phase-A boxes: [mid - half, mid + half] must contain the padded box in real arithmetic, so `half` is computed in
        // double from the rounded float `mid` and then rounded up (two float steps cover the double rounding as well)
        const float nan = std::numeric_limits<float>::quiet_NaN();
        std::vector<float> mid(3 * TRT_FLAT_MAX, nan), half(3 * TRT_FLAT_MAX, nan);
        double bound = 0.0; bool finite = true;
        for (size_t i = 0; i < n; ++i) {
            const HostPrim& p = prims[i];
            const size_t r = leaf_rank[i];
            for (int k = 0; k < 3; ++k) {
                const double lo = p.bmin[k], hi = p.bmax[k];
                const float m = (float)(0.5 * (lo + hi));
                float h = (float)std::max((double)m - lo, hi - (double)m);
                h = std::nextafter(std::nextafter(h, std::numeric_limits<float>::infinity()), std::numeric_limits<float>::infinity());
                mid[3 * r + k] = m; half[3 * r + k] = h;
                if (!std::isfinite(lo) || !std::isfinite(hi) || !std::isfinite(h)) finite = false;
                bound = std::max(bound, std::max(std::fabs(lo), std::fabs(hi)));
            }
        }
        flat.flat_cbox.resize(3 * TRT_FLAT_MAX / 2);
        for (size_t j = 0; j < TRT_FLAT_MAX / 2; ++j) {
            const float *m0 = &mid[6 * j], *m1 = &mid[6 * j + 3], *h0 = &half[6 * j], *h1 = &half[6 * j + 3];
            flat.flat_cbox[3 * j] = {m0[0], m1[0], m0[1], m1[1]};
            flat.flat_cbox[3 * j + 1] = {m0[2], m1[2], h0[0], h1[0]};
            flat.flat_cbox[3 * j + 2] = {h0[1], h1[1], h0[2], h1[2]};
        }
        flat.flat_bound = (finite && bound <= 1073741824.0) ? (float)bound : -1.f;
    }
    flat.materials.resize(2 * materials.size());
    for (size_t m = 0; m < materials.size(); ++m) {
        flat.materials[2 * m] = {materials[m].rgb[0], materials[m].rgb[1], materials[m].rgb[2], as_float(materials[m].kind)};
        flat.materials[2 * m + 1] = {materials[m].param, 0.f, 0.f, 0.f};
    }
    committed = true; version++;
    return true;
}

} // namespace trt
