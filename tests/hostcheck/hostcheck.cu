// hostcheck — TEST INFRASTRUCTURE.  Compiles the device headers of cgal_b200/csrc for the HOST so that the
// per-thread logic of the CUDA kernels (predicates, constructions, Karras topology, Morton keys, traversal) can be
// checked against the oracle without a GPU (pytest -m "not gpu").  It is not part of libb200aabb.so and is never
// used as a fallback: the product library has no host execution path.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

static std::vector<uint32_t>* g_node_trace = nullptr;       // nodes fetched by traverse_closest (hc_closest_depth_histogram)
static void trace_node(uint32_t index) { if (g_node_trace) g_node_trace->push_back(index); }
#if defined(__CUDA_ARCH__)
#define B2_TRACE_NODE(index)
#else
#define B2_TRACE_NODE(index) trace_node(index)
#endif
#include "../../cgal_b200/csrc/build.cuh"
#include "../../cgal_b200/csrc/closest_warp.cuh"
#include "../../cgal_b200/csrc/ingest.cuh"
#define B2_WARP_EMU 1
#include "warp_emu.h"

using namespace b2;

namespace {
struct HostTree { std::vector<BvhNode> nodes; std::vector<TriRec> tris; uint32_t root = kNone; uint64_t n = 0; double bbox[6]; std::vector<uint64_t> keys;
    double mag() const { double m = 0; for (int a = 0; a < 6; ++a) if (std::isfinite(bbox[a])) m = std::max(m, std::fabs(bbox[a])); return m; } };
Vec3 V(const double* p) { return {p[0], p[1], p[2]}; }
Tri3 T(const double* p) { return {V(p), V(p + 3), V(p + 6)}; }

void refit(HostTree& h, uint32_t ref, float* lo, float* hi) {
    if (ref & kLeafBit) {
        const double* v = h.tris[ref & ~kLeafBit].v;
        for (int a = 0; a < 3; ++a) { lo[a] = float_below(std::min(v[a], std::min(v[3 + a], v[6 + a]))); hi[a] = float_above(std::max(v[a], std::max(v[3 + a], v[6 + a]))); }
        return;
    }
    BvhNode& nd = h.nodes[ref];
    refit(h, nd.child0, nd.lo0, nd.hi0);
    refit(h, nd.child1, nd.lo1, nd.hi1);
    for (int a = 0; a < 3; ++a) { lo[a] = std::min(nd.lo0[a], nd.lo1[a]); hi[a] = std::max(nd.hi0[a], nd.hi1[a]); }
}
}  // namespace

extern "C" {

// emulates b200aabb_build step by step with the same per-element device functions
void* hc_build(const double* tris, const uint32_t* ids, uint64_t n) {
    HostTree* h = new HostTree();
    h->n = n;
    for (int a = 0; a < 3; ++a) { h->bbox[a] = INFINITY; h->bbox[3 + a] = -INFINITY; }
    if (n == 0) return h;
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint64_t i = 0; i < n; ++i) for (int k = 0; k < 9; ++k) { lo[k % 3] = std::min(lo[k % 3], tris[9 * i + k]); hi[k % 3] = std::max(hi[k % 3], tris[9 * i + k]); }
    double scale[3], lo_[3];
    double weight[3] = {1.0, 1.0, 1.0};
    if (const char* e = getenv("HC_KEY_WEIGHTS")) sscanf(e, "%lf,%lf,%lf", &weight[0], &weight[1], &weight[2]);     // experiment hook (scripts/key_weight_experiment.py)
    morton_frame(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], lo_, scale, weight);
    std::vector<uint64_t> keys(n);
    for (uint64_t i = 0; i < n; ++i) {
        double c[3];
        for (int k = 0; k < 3; ++k) { const double a = tris[9 * i + k], b = tris[9 * i + 3 + k], d = tris[9 * i + 6 + k]; c[k] = 0.5 * (fmin(a, fmin(b, d)) + fmax(a, fmax(b, d))); }
        keys[i] = morton63(c[0], c[1], c[2], lo, scale);
    }
    // the build's sort words (build.cuh::pack_sort_word), with its fall-back to full (key, index) pairs when many primitives share their kept key bits
    const int ib = index_bits_for(n);
    std::vector<uint64_t> words(n);
    for (uint64_t i = 0; i < n; ++i) words[i] = pack_sort_word(keys[i], i, ib);
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return words[a] < words[b]; });
    uint64_t dups = 0;
    for (uint64_t k = 0; k + 3 < n; ++k) dups += (((words[perm[k]] ^ words[perm[k + 1]]) | (words[perm[k]] ^ words[perm[k + 2]]) | (words[perm[k]] ^ words[perm[k + 3]])) >> ib) == 0 ? 1 : 0;
    const bool packed = dups <= n / 16;
    if (!packed) { std::iota(perm.begin(), perm.end(), 0u); std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; }); }
    h->keys.resize(n);
    h->tris.resize(n);
    for (uint64_t k = 0; k < n; ++k) {
        h->keys[k] = packed ? words[perm[k]] : keys[perm[k]];
        TriRec r; memset(&r, 0, sizeof(r));
        memcpy(r.v, tris + 9 * (uint64_t)perm[k], 72);
        r.id = ids ? ids[perm[k]] : perm[k];
        r.pad[0] = perm[k];
        h->tris[k] = r;
    }
    for (int a = 0; a < 3; ++a) { h->bbox[a] = lo[a]; h->bbox[3 + a] = hi[a]; }     // bbox() = the exact scene bounds
    if (n == 1) { h->root = kLeafBit; return h; }
    h->nodes.resize(n - 1);
    memset(h->nodes.data(), 0, (n - 1) * sizeof(BvhNode));
    h->nodes[0].parent = kNone;
    for (uint64_t i = 0; i + 1 < n; ++i) {
        uint32_t l, r;
        karras_node(h->keys.data(), (int64_t)n, (int64_t)i, l, r);
        h->nodes[i].child0 = l; h->nodes[i].child1 = r;
        if (!(l & kLeafBit)) h->nodes[l].parent = (uint32_t)i;
        if (!(r & kLeafBit)) h->nodes[r].parent = (uint32_t)i;
    }
    h->root = 0;
    float rlo[3], rhi[3];
    refit(*h, 0, rlo, rhi);
    return h;
}
// Re-builds the nodes of a tree with the FUSED bottom-up path of b200aabb_build (build.cuh: build_local_climb per block of kRefitBlock leaves with the
// threads run one after the other, then Karras' searches for the nodes not built there, then the frontier climbs) and returns the number of node words
// (boxes, children, parent, range end) that differ from the classic emulation of hc_build: 0 = identical.  stats[0] = frontier entries, [1] = nodes left to
// the searches, [2] = frontier overflow.  frontier_per_block < 0: the build's own capacity.
int64_t hc_fused_build_diff(void* hv, int frontier_per_block, uint64_t* stats) {
    HostTree* h = (HostTree*)hv;
    const int64_t n = (int64_t)h->n;
    stats[0] = stats[1] = stats[2] = 0;
    if (n < 2) return 0;
    std::vector<BvhNode> nodes(n - 1);
    memset(nodes.data(), 0xCD, (n - 1) * sizeof(BvhNode));
    std::vector<uint32_t> leaf_parent(n, 0xCDCDCDCDu);
    const int64_t blocks = (n + kRefitBlock - 1) / kRefitBlock;
    const unsigned cap = (unsigned)(blocks * (frontier_per_block < 0 ? kRefitFrontierPerBlock : frontier_per_block));
    std::vector<RefitEntry> frontier(std::max<unsigned>(cap, 1));
    unsigned count = 0, overflow = 0;
    std::vector<uint32_t> done(blocks * (kRefitBlock / 32), 0);
    for (int64_t b = 0; b < blocks; ++b) {
        LocalBuildShared S;
        memset(&S, 0, sizeof(S));
        const int64_t base = b * kRefitBlock, last_leaf = std::min<int64_t>(base + kRefitBlock - 1, n - 1);
        for (int64_t k = base; k <= last_leaf; ++k) S.key[k - base + 1] = h->keys[k];
        S.key[0] = base > 0 ? h->keys[base - 1] : 0;
        S.key[last_leaf - base + 2] = last_leaf + 1 < n ? h->keys[last_leaf + 1] : 0;
        for (int64_t k = base; k <= last_leaf; ++k) S.len[k - base + 1] = local_len_entry(S, base, last_leaf, n, k);
        S.len[0] = local_len_entry(S, base, last_leaf, n, base - 1);
        for (int64_t k = base; k <= last_leaf; ++k) {
            const double* v = h->tris[k].v;
            float lo[3], hi[3];
            for (int a = 0; a < 3; ++a) { lo[a] = float_below(std::min(v[a], std::min(v[3 + a], v[6 + a]))); hi[a] = float_above(std::max(v[a], std::max(v[3 + a], v[6 + a]))); }
            build_local_climb(S, base, last_leaf, n, k, lo, hi, nodes.data(), leaf_parent.data(), frontier.data(), &count, cap, &overflow);
        }
        for (int64_t k = base; k <= last_leaf; ++k) if (S.done[k - base]) done[k >> 5] |= 1u << (k & 31);
    }
    stats[0] = count; stats[2] = overflow;
    if (overflow) return -1;
    for (int64_t i = 0; i + 1 < n; ++i) {                 // hierarchy_kernel with the `done` bitmap
        if ((done[i >> 5] >> (i & 31)) & 1u) continue;
        ++stats[1];
        uint32_t l, r, j;
        karras_node(h->keys.data(), n, i, l, r, &j);
        nodes[i].child0 = l; nodes[i].child1 = r; nodes[i].visit = j;
        if (i == 0) nodes[i].parent = kNone;
        if (l & kLeafBit) leaf_parent[l & ~kLeafBit] = (uint32_t)i; else nodes[l].parent = (uint32_t)i;
        if (r & kLeafBit) leaf_parent[r & ~kLeafBit] = (uint32_t)i; else nodes[r].parent = (uint32_t)i;
    }
    for (unsigned e = 0; e < count; ++e) {                // refit_global_kernel, one entry after the other
        float lo[3] = {frontier[e].lo[0], frontier[e].lo[1], frontier[e].lo[2]}, hi[3] = {frontier[e].hi[0], frontier[e].hi[1], frontier[e].hi[2]};
        uint32_t me = frontier[e].from;
        uint32_t p = (me & kLeafBit) ? leaf_parent[me & ~kLeafBit] : nodes[me].parent;
        while (p != kNone) {
            BvhNode& nd = nodes[p];
            const int side = nd.child0 == me ? 0 : 1;
            float* mylo = side ? nd.lo1 : nd.lo0; float* myhi = side ? nd.hi1 : nd.hi0;
            for (int a = 0; a < 3; ++a) { mylo[a] = lo[a]; myhi[a] = hi[a]; }
            const uint32_t old = nd.visit;
            nd.visit ^= 0x80000000u;
            if ((old & 0x80000000u) == 0) break;
            const float* olo = side ? nd.lo0 : nd.lo1; const float* ohi = side ? nd.hi0 : nd.hi1;
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], olo[a]); hi[a] = std::max(hi[a], ohi[a]); }
            me = p; p = nd.parent;
        }
    }
    // the classic emulation does not fill `visit`: compare it with the range end Karras' search gives
    int64_t diff = 0;
    for (int64_t i = 0; i + 1 < n; ++i) {
        const uint32_t* a = (const uint32_t*)&nodes[i]; const uint32_t* b = (const uint32_t*)&h->nodes[i];
        for (int w = 0; w < 15; ++w) diff += a[w] != b[w];
        uint32_t l, r, j;
        karras_node(h->keys.data(), n, i, l, r, &j);
        diff += nodes[i].visit != j;
    }
    for (int64_t k = 0; k < n; ++k) {
        const uint32_t p = leaf_parent[k];
        diff += !(p < (uint32_t)(n - 1) && (h->nodes[p].child0 == (kLeafBit | (uint32_t)k) || h->nodes[p].child1 == (kLeafBit | (uint32_t)k)));
    }
    return diff;
}
void hc_destroy(void* hv) { delete (HostTree*)hv; }
uint64_t hc_num_nodes(void* hv) { return ((HostTree*)hv)->nodes.size(); }
void hc_export(void* hv, void* nodes, void* tris, double* bbox) {
    HostTree* h = (HostTree*)hv;
    if (nodes) memcpy(nodes, h->nodes.data(), h->nodes.size() * sizeof(BvhNode));
    if (tris) memcpy(tris, h->tris.data(), h->tris.size() * sizeof(TriRec));
    if (bbox) memcpy(bbox, h->bbox, 48);
}
// structural validation of a node/leaf array (used on the GPU-built tree too): every leaf reachable exactly once,
// every child box bit-equal to the outward binary32 rounding of the exact union of the boxes below it.  returns 0 if valid.
int hc_validate(const void* nodes_v, const void* tris_v, uint64_t n) {
    if (n < 2) return 0;
    const BvhNode* nodes = (const BvhNode*)nodes_v; const TriRec* tris = (const TriRec*)tris_v;
    std::vector<uint8_t> seen(n, 0), nseen(n - 1, 0);
    struct Item { uint32_t ref; };
    // iterative post-order computing boxes
    std::vector<float> blo(3 * (n - 1)), bhi(3 * (n - 1));
    std::vector<uint32_t> order; order.reserve(n - 1);
    std::vector<uint32_t> st{0};
    while (!st.empty()) {
        uint32_t x = st.back(); st.pop_back();
        if (x >= n - 1 || nseen[x]) return 1;
        nseen[x] = 1; order.push_back(x);
        for (int c = 0; c < 2; ++c) { uint32_t r = nodes[x].child(c); if (r & kLeafBit) { uint32_t s = r & ~kLeafBit; if (s >= n || seen[s]) return 2; seen[s] = 1; } else { if (nodes[r].parent != x) return 3; st.push_back(r); } }
    }
    for (uint64_t i = 0; i < n; ++i) if (!seen[i]) return 4;
    for (size_t k = order.size(); k-- > 0;) {
        const uint32_t x = order[k];
        for (int c = 0; c < 2; ++c) {
            float lo[3], hi[3];
            const uint32_t r = nodes[x].child(c);
            if (r & kLeafBit) { const double* v = tris[r & ~kLeafBit].v; for (int a = 0; a < 3; ++a) { lo[a] = float_below(std::min(v[a], std::min(v[3 + a], v[6 + a]))); hi[a] = float_above(std::max(v[a], std::max(v[3 + a], v[6 + a]))); } }
            else for (int a = 0; a < 3; ++a) { lo[a] = blo[3 * r + a]; hi[a] = bhi[3 * r + a]; }
            const float* nlo = c ? nodes[x].lo1 : nodes[x].lo0; const float* nhi = c ? nodes[x].hi1 : nodes[x].hi0;
            for (int a = 0; a < 3; ++a) if (nlo[a] != lo[a] || nhi[a] != hi[a]) return 5;
            for (int a = 0; a < 3; ++a) { if (c == 0) { blo[3 * x + a] = lo[a]; bhi[3 * x + a] = hi[a]; } else { blo[3 * x + a] = std::min(blo[3 * x + a], lo[a]); bhi[3 * x + a] = std::max(bhi[3 * x + a], hi[a]); } }
        }
    }
    return 0;
}

void hc_intersect(void* hv, const double* q, uint64_t n, int is_ray, int count, uint32_t* out) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        if (h->n == 0) { out[i] = 0; continue; }
        const Vec3 s = V(q + 6 * i), e = V(q + 6 * i + 3);
        if (is_ray) out[i] = count ? traverse_intersect<kRayQ, kCount>(h->nodes.data(), h->tris.data(), h->root, s, e, st) : traverse_intersect<kRayQ, kAnyHit>(h->nodes.data(), h->tris.data(), h->root, s, e, st);
        else out[i] = count ? traverse_intersect<kSegmentQ, kCount>(h->nodes.data(), h->tris.data(), h->root, s, e, st) : traverse_intersect<kSegmentQ, kAnyHit>(h->nodes.data(), h->tris.data(), h->root, s, e, st);
    }
}
// Line_3 (kind 2: point + direction vector, 6 doubles) and Plane_3 (kind 3: a, b, c, d) counts through the device traversal routines
void hc_count_line_plane(void* hv, const double* q, uint64_t n, int kind, uint32_t* out) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        if (h->n == 0) { out[i] = 0; continue; }
        if (kind == 2) out[i] = traverse_intersect<kLineQ, kCount>(h->nodes.data(), h->tris.data(), h->root, V(q + 6 * i), V(q + 6 * i + 3), st);
        else out[i] = traverse_plane<kCount>(h->nodes.data(), h->tris.data(), h->root, q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3], st);
    }
}
// how many of the nodes a closest_point query fetches lie at each depth of the tree (root = 0): hist[64] summed over the queries; returns the tree depth
int hc_closest_depth_histogram(void* hv, const double* q, uint64_t n, uint64_t* hist) {
    HostTree* h = (HostTree*)hv;
    std::vector<uint8_t> depth(h->nodes.size(), 0);
    std::vector<uint32_t> order{h->root};
    int maxd = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        const BvhNode& nd = h->nodes[order[i]];
        for (uint32_t c : {nd.child0, nd.child1})
            if (!(c & kLeafBit)) { depth[c] = (uint8_t)std::min(63, depth[order[i]] + 1); maxd = std::max<int>(maxd, depth[c]); order.push_back(c); }
    }
    std::vector<uint32_t> trace;
    g_node_trace = &trace;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        trace.clear();
        traverse_closest(h->nodes.data(), h->tris.data(), h->root, V(q + 3 * i), h->mag(), st);
        for (uint32_t nd : trace) ++hist[depth[nd]];
    }
    g_node_trace = nullptr;
    return maxd;
}
void hc_first_intersection(void* hv, const double* rays, uint64_t n, uint32_t* prim, double* t, double* pts, uint8_t* kind, uint8_t* tie) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        if (h->n == 0) { prim[i] = kNone; t[i] = 0; kind[i] = 0; tie[i] = 0; memset(pts + 6 * i, 0, 48); continue; }
        const FirstHit f = traverse_first(h->nodes.data(), h->tris.data(), h->root, V(rays + 6 * i), V(rays + 6 * i + 3), h->mag(), st);
        prim[i] = f.prim; t[i] = f.prim == kNone ? 0.0 : f.t; kind[i] = (uint8_t)f.kind; tie[i] = (uint8_t)(f.tie | (f.near << 1));
        memset(pts + 6 * i, 0, 48);
        if (f.prim != kNone) { pts[6 * i] = f.p0.x; pts[6 * i + 1] = f.p0.y; pts[6 * i + 2] = f.p0.z; if (f.kind == 2) { pts[6 * i + 3] = f.p1.x; pts[6 * i + 4] = f.p1.y; pts[6 * i + 5] = f.p1.z; } }
    }
}
void hc_closest_point(void* hv, const double* q, uint64_t n, double* cp, uint32_t* prim, double* sqd, uint8_t* tie) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        const Closest c = traverse_closest(h->nodes.data(), h->tris.data(), h->root, V(q + 3 * i), h->mag(), st);
        cp[3 * i] = c.p.x; cp[3 * i + 1] = c.p.y; cp[3 * i + 2] = c.p.z; prim[i] = c.prim; sqd[i] = c.sqd; tie[i] = (uint8_t)(c.tie | (c.near << 1));
    }
}

// a tree over Segment_3 primitives stored as (source, target, target) records: bvh.cuh::traverse_closest<true> and traverse_triangle<.., true>
void hc_closest_point_segments(void* hv, const double* q, uint64_t n, double* cp, uint32_t* prim, double* sqd, uint8_t* tie) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        const Closest c = traverse_closest<true>(h->nodes.data(), h->tris.data(), h->root, V(q + 3 * i), h->mag(), st);
        cp[3 * i] = c.p.x; cp[3 * i + 1] = c.p.y; cp[3 * i + 2] = c.p.z; prim[i] = c.prim; sqd[i] = c.sqd; tie[i] = (uint8_t)(c.tie | (c.near << 1));
    }
}
void hc_count_triangle_segments(void* hv, const double* q, uint64_t n, uint32_t* out) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        out[i] = h->n == 0 ? 0 : traverse_triangle<kCount, NoXf, true>(h->nodes.data(), h->tris.data(), h->root, T(q + 9 * i), st);
    }
}

// ingest.cuh::parse_decimal on a batch of literals packed in one buffer (offsets[n + 1]): bits + status per literal
void hc_parse_decimals(const char* buf, const uint64_t* offsets, uint64_t n, uint64_t* bits, int* status) {
    for (uint64_t i = 0; i < n; ++i) bits[i] = parse_decimal(buf + offsets[i], buf + offsets[i + 1], status[i]);
}

// predicate / construction level
int hc_orient3(const double* p) { int n = 0; return orient3(V(p), V(p + 3), V(p + 6), V(p + 9), n); }
int hc_orient3_exact(const double* p) {
    return exact_det3_sign(dd_diff(p[3], p[0]), dd_diff(p[4], p[1]), dd_diff(p[5], p[2]), dd_diff(p[6], p[0]), dd_diff(p[7], p[1]), dd_diff(p[8], p[2]),
                           dd_diff(p[9], p[0]), dd_diff(p[10], p[1]), dd_diff(p[11], p[2]));
}
int hc_orient3_dir(const double* p) { int n = 0; return orient3_dir(V(p), V(p + 3), V(p + 6), V(p + 9), V(p + 12), n); }
int hc_coplanar_orient(const double* p) { int n = 0; return coplanar_orient(V(p), V(p + 3), V(p + 6), n); }
int hc_cmp_dist(const double* p) { int n = 0; return cmp_dist(V(p), V(p + 3), V(p + 6), n); }
int hc_cmp_dist_exact(const double* p) { return exact_cmp_dist_sign(p, p + 3, p + 6); }
int hc_classify_ray_tri(const double* tri, const double* pq) { int n = 0; return ray_tri_classify(T(tri), V(pq), V(pq + 3), n); }
// Side_of_triangle_mesh through bvh.cuh::traverse_side: rays along the given directions (ndir x 3) until one is decisive; -2 = none was
void hc_side_of_mesh(void* hv, const double* pts, uint64_t n, const double* dirs, int ndir, int8_t* side, uint8_t* attempts) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        const Vec3 p = V(pts + 3 * i);
        side[i] = -1; attempts[i] = 0;
        if (h->n == 0 || p.x < h->bbox[0] || p.x > h->bbox[3] || p.y < h->bbox[1] || p.y > h->bbox[4] || p.z < h->bbox[2] || p.z > h->bbox[5]) continue;
        int r = kSideIndeterminate;
        for (int k = 0; k < ndir && r == kSideIndeterminate; ++k) {
            TravStat st;
            const Vec3 q = {p.x + dirs[3 * k], p.y + dirs[3 * k + 1], p.z + dirs[3 * k + 2]};
            r = traverse_side(h->nodes.data(), h->tris.data(), h->root, p, q, st);
            ++attempts[i];
        }
        side[i] = (int8_t)(r == kSideIndeterminate ? -2 : (r == kSideBoundary ? 0 : ((r & 1) ? 1 : -1)));
    }
}
// the same with the mesh under the affine map m (12 doubles): no root-box shortcut, the traversal decides (a point far outside crosses nothing)
void hc_side_of_mesh_xf(void* hv, const double* m, const double* pts, uint64_t n, const double* dirs, int ndir, int8_t* side) {
    HostTree* h = (HostTree*)hv;
    AffXf x; for (int i = 0; i < 12; ++i) x.m[i] = m[i];
    for (uint64_t i = 0; i < n; ++i) {
        const Vec3 p = V(pts + 3 * i);
        side[i] = -1;
        if (h->n == 0) continue;
        int r = kSideIndeterminate;
        for (int k = 0; k < ndir && r == kSideIndeterminate; ++k) {
            TravStat st;
            const Vec3 q = {p.x + dirs[3 * k], p.y + dirs[3 * k + 1], p.z + dirs[3 * k + 2]};
            r = traverse_side<AffXf>(h->nodes.data(), h->tris.data(), h->root, p, q, st, x);
        }
        side[i] = (int8_t)(r == kSideIndeterminate ? -2 : (r == kSideBoundary ? 0 : ((r & 1) ? 1 : -1)));
    }
}
// tree.do_intersect(other_tree) with both trees under affine maps, as b200aabb.cu::tree_tree_kernel<true> runs it: every primitive of `small`, moved by mq,
// traverses `big` under mb (bvh.cuh::traverse_triangle<kAnyId, AffXf>); returns 1 + the ids of the first pair found, 0 if none
int hc_tree_tree_xf(void* big, const double* mb, void* small, const double* mq, uint32_t* pair) {
    HostTree* hb = (HostTree*)big; HostTree* hs = (HostTree*)small;
    AffXf xq, xb;
    for (int i = 0; i < 12; ++i) { xq.m[i] = mq[i]; xb.m[i] = mb[i]; }
    if (hb->n == 0) return 0;
    for (uint64_t i = 0; i < hs->n; ++i) {
        TravStat st;
        uint32_t ida;
        const Tri3 t = xq.apply(load_tri(hs->tris.data(), (uint32_t)i, ida));
        const uint32_t idb = traverse_triangle<kAnyId, AffXf>(hb->nodes.data(), hb->tris.data(), hb->root, t, st, nullptr, 0, xb);
        if (idb != kNone) { pair[0] = ida; pair[1] = idb; return 1; }
    }
    return 0;
}
int hc_tri_tri_do_intersect(const double* query, const double* tri) { int n = 0; return tri_tri_do_intersect(T(query), T(tri), n) ? 1 : 0; }
// Triangle_3 queries (9 doubles each) through the device traversal routine: mode 0 do_intersect, 1 count
void hc_intersect_triangle(void* hv, const double* q, uint64_t n, int count, uint32_t* out) {
    HostTree* h = (HostTree*)hv;
    for (uint64_t i = 0; i < n; ++i) {
        TravStat st;
        if (h->n == 0) { out[i] = 0; continue; }
        const Tri3 t = T(q + 9 * i);
        out[i] = count ? traverse_triangle<kCount>(h->nodes.data(), h->tris.data(), h->root, t, st) : traverse_triangle<kAnyHit>(h->nodes.data(), h->tris.data(), h->root, t, st);
    }
}
// bvh.cuh::AffXf: the image of a triangle under a 3 x 4 affine map, and the conservative node test "transformed box [box] (binary32 bounds) meets the box of query"
void hc_xf_apply(const double* m, const double* tri, double* out) {
    AffXf x; for (int i = 0; i < 12; ++i) x.m[i] = m[i];
    const Tri3 t = x.apply(T(tri));
    out[0] = t.a.x; out[1] = t.a.y; out[2] = t.a.z; out[3] = t.b.x; out[4] = t.b.y; out[5] = t.b.z; out[6] = t.c.x; out[7] = t.c.y; out[8] = t.c.z;
}
int hc_xf_box_meets(const double* m, const float* box, const double* query) {
    AffXf x; for (int i = 0; i < 12; ++i) x.m[i] = m[i];
    return x.box_meets(box, box + 3, make_box_of(T(query))) ? 1 : 0;
}
int hc_tri_do_intersect_kind(const double* tri, const double* q, int kind) {
    int n = 0;
    if (kind == 2) return line_tri_do_intersect(T(tri), V(q), V(q + 3), n);
    if (kind == 3) return plane_tri_do_intersect(T(tri), q[0], q[1], q[2], q[3], n);
    return kind == 1 ? ray_tri_do_intersect(T(tri), V(q), V(q + 3), n) : seg_tri_do_intersect(T(tri), V(q), V(q + 3), n);
}
int hc_tri_do_intersect(const double* tri, const double* pq, int is_ray) { int n = 0; return is_ray ? ray_tri_do_intersect(T(tri), V(pq), V(pq + 3), n) : seg_tri_do_intersect(T(tri), V(pq), V(pq + 3), n); }
int hc_tri_query_intersection(const double* tri, const double* q, int kind, double* pts) {      // kind 2 Line_3 (point, vector), 3 Plane_3 (a, b, c, d)
    int n = 0;
    const SegX x = kind == 2 ? line_tri_intersection(T(tri), V(q), V(q + 3), n) : plane_tri_intersection(T(tri), q[0], q[1], q[2], q[3], n);
    if (x.kind == 1 || x.kind == 2) { pts[0] = x.p0.x; pts[1] = x.p0.y; pts[2] = x.p0.z; if (x.kind == 2) { pts[3] = x.p1.x; pts[4] = x.p1.y; pts[5] = x.p1.z; } }
    return x.kind;
}
int hc_tri_seg_intersection(const double* tri, const double* seg, double* pts) {
    int n = 0;
    const SegX x = seg_tri_intersection(T(tri), V(seg), V(seg + 3), n);
    if (x.kind) { pts[0] = x.p0.x; pts[1] = x.p0.y; pts[2] = x.p0.z; if (x.kind == 2) { pts[3] = x.p1.x; pts[4] = x.p1.y; pts[5] = x.p1.z; } }
    return x.kind;
}
int hc_tri_ray_intersection(const double* tri, const double* ray, double* pts, double* t) {
    int n = 0; const Vec3 s = V(ray), q = V(ray + 3), d = q - s;
    const RayX x = ray_tri_intersection(T(tri), s, q, d, ray_major_axis(d), n);
    if (x.kind) { pts[0] = x.p0.x; pts[1] = x.p0.y; pts[2] = x.p0.z; pts[3] = x.p1.x; pts[4] = x.p1.y; pts[5] = x.p1.z; *t = x.t; }
    return x.kind;
}
void hc_closest_on_triangle(const double* tri, const double* q, double* out) { const Vec3 c = closest_on_triangle(T(tri), V(q)); out[0] = c.x; out[1] = c.y; out[2] = c.z; }
// the box is rounded outward to binary32 exactly as the refit does
int hc_line_box(const double* box, const double* pq, int is_ray) {
    const LineCtx c = make_line(V(pq), V(pq + 3));
    float lo[3], hi[3], tn;
    for (int a = 0; a < 3; ++a) { lo[a] = float_below(box[a]); hi[a] = float_above(box[3 + a]); }
    return line_box(lo, hi, c, is_ray ? INFINITY : 1.0f, tn) ? 1 : 0;
}
// lower bound of the squared distance point - box computed by the traversal (binary32, rounded down)
double hc_point_box_sqdist(const double* box, const double* p) {
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = float_below(box[a]); hi[a] = float_above(box[3 + a]); }
    return (double)point_box_sqdist(lo, hi, make_point(V(p)));
}
// The binary32 distance bounds of the warp-cooperative closest_point kernel (closest_warp.cuh), computed exactly as its pre-test
// stages compute them for one (query, primitive) pair: out = {lower bound of stage 1, lower bound of stage 2, upper bound of stage 2
// (NaN = none), e2, dslack}.  mag = largest coordinate magnitude of the scene.
void hc_warp_bounds(const double* tri, const double* q, double mag, float* out) {
    const float qx = (float)q[0], qy = (float)q[1], qz = (float)q[2];
    const double ex = q[0] - (double)qx, ey = q[1] - (double)qy, ez = q[2] - (double)qz;
    const float perr = float_above(std::sqrt((ex * ex + ey * ey) + ez * ez) * 1.0000001) + 1e-37f;
    const float ev = float_above(mag * 5.9604645e-8 * 1.000001);
    const float e2 = float_above((double)float_above(1.7320509 * (double)ev) + (double)perr);        // >= fadd_ru(fmul_ru(sqrt3, ev), perr)
    const double delta = 4e-16 * std::max(mag, std::max(std::fabs(q[0]), std::max(std::fabs(q[1]), std::fabs(q[2]))));
    const float dslack = float_above(2.0 * delta) + 1e-37f;
    float f[9];
    for (int k = 0; k < 9; ++k) f[k] = (float)tri[k];                                                  // the TriF record
    RelTri r;                                                                                        // as wp_relative
    r.ra[0] = f[0] - qx; r.ra[1] = f[1] - qy; r.ra[2] = f[2] - qz;
    r.rb[0] = f[3] - qx; r.rb[1] = f[4] - qy; r.rb[2] = f[5] - qz;
    r.rc[0] = f[6] - qx; r.rc[1] = f[7] - qy; r.rc[2] = f[8] - qz;
    r.m = 0.0f;
    for (int k = 0; k < 3; ++k) r.m = std::max(r.m, std::max(std::fabs(r.ra[k]), std::max(std::fabs(r.rb[k]), std::fabs(r.rc[k]))));
    const float g[3] = {(r.ra[0] + r.rb[0]) + r.rc[0], (r.ra[1] + r.rb[1]) + r.rc[1], (r.ra[2] + r.rb[2]) + r.rc[2]};
    out[0] = wp_lower(g, sqrtf(dot3f(g, g)), r) - e2;
    wp_sharp(r, dslack, e2, &out[1], &out[2]);
    out[3] = e2; out[4] = dslack;
}
// closest_pair.cuh's kernel body run by `warps` emulated warps (32 fibers each, one warp after the other sharing the batch cursor) over
// the host-built tree: the SAME code the GPU runs.  perm = optional coherence order.  Returns 0, or 1 when a warp dead-locked (non-uniform
// control flow at a collective), 2 when the kernel reported an overflow.  stats[0..2] = nodes visited, leaves tested, leaves culled.
int hc_closest_point_pair(void* hv, const double* q, uint64_t n, const uint32_t* perm, int warps, int exact_go, double* cp, uint32_t* prim, double* sqd,
                          uint8_t* flags, uint64_t* stats) {
    HostTree* h = (HostTree*)hv;
    std::vector<TriF> trisf(h->tris.size());
    for (size_t k = 0; k < h->tris.size(); ++k) { memset(&trisf[k], 0, sizeof(TriF)); for (int j = 0; j < 9; ++j) trisf[k].v[j] = (float)h->tris[k].v[j]; }
    unsigned long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cursor = 0;
    std::vector<uint2> spill((size_t)warps * kPpSpillPerWarp);
    std::vector<WarpSlots64> slots((size_t)warps);
    PairClosestArgs a;
    a.nodes = h->nodes.data(); a.tris = h->tris.data(); a.trisf = trisf.data(); a.mag = h->mag(); a.perm = perm; a.counters = counters;
    a.pts = q; a.n = n; a.cp = cp; a.prim = prim; a.sqd = sqd; a.flags = flags; a.cursor = &cursor; a.spill = spill.data(); a.slots64 = slots.data();
    a.exact_go = exact_go; a.flag_tie = 4; a.flag_exact = 8;
    // the warps run one after the other (each to completion): warp 0 would take every chunk, so the cursor is advanced by hand to deal the
    // chunks round-robin — warp k sees a cursor that skips the other warps' chunks
    int rc = 0;
    for (int k = 0; k < warps && rc == 0; ++k) {
        std::vector<unsigned long long> handed;
        PairShared S;
        memset(&S, 0, sizeof(S));
        b2emu::Warp W;
        // a private view of the batch: queries whose chunk index % warps == k
        std::vector<uint32_t> myperm;
        for (uint64_t c = (uint64_t)k * kPpChunk; c < n; c += (uint64_t)warps * kPpChunk)
            for (uint64_t i = c; i < std::min<uint64_t>(n, c + kPpChunk); ++i) myperm.push_back(perm ? perm[i] : (uint32_t)i);
        PairClosestArgs b = a;
        unsigned long long mycursor = 0;
        b.perm = myperm.data(); b.n = myperm.size(); b.cursor = &mycursor;
        b.spill = spill.data() + (size_t)k * kPpSpillPerWarp;
        if (b.n == 0) continue;
        const bool ok = b2emu::run_warp(W, [&](unsigned lane) {
            WarpCtx w; w.lane = lane; w.emu = &W; w.phase = 0;
            closest_pair_body(b, S, slots[k], b.spill, w);
        });
        if (!ok) rc = 1;
    }
    if (rc == 0 && counters[2]) rc = 2;
    if (stats) { stats[0] = counters[3]; stats[1] = counters[4]; stats[2] = counters[6]; }
    return rc;
}
uint64_t hc_sizeof_node() { return sizeof(BvhNode); }
uint64_t hc_sizeof_tri() { return sizeof(TriRec); }

}  // extern "C"
