// b200aabb.cu — libb200aabb.so: traversal kernels + the C ABI declared in include/b200aabb.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// (-fmad=false: the constructions of geom.cuh must not be contracted into FMAs; explicit fma() calls remain FMAs).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200aabb.h"
#include "build.cuh"
#include "closest_warp.cuh"
#include "ingest.cuh"
#include "closest_pair.cuh"

using namespace b2;

// ------------------------------------------------------------------------------------------------------------------
// Query kernels: one thread per query, iterative stack traversal (bvh.cuh).
// counters[0] += queries that used the exact stage, counters[1] += ties, counters[2] |= stack overflow
// ------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ Vec3 ld3(const double* __restrict__ p) { return {__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }

// what every traversal kernel needs to know about the tree
struct TreeRef {
    const BvhNode* nodes;
    const TriRec* tris;
    uint32_t root;
    double mag;                       // largest coordinate magnitude of the root box
    const uint32_t* perm;             // coherence order of the batch (thread i serves query perm[i]); nullptr = identity
    unsigned long long* counters;     // [8]: exact-stage queries, ties, stack overflow, nodes visited, leaves visited, near ties, leaves culled by the binary32 pre-test
};

// one atomic per warp and counter instead of one per thread
__device__ __forceinline__ void flush_stats(const TreeRef& tr, const TravStat& st, int tie, int near) {
    const unsigned full = 0xFFFFFFFFu;
    const unsigned nodes = __reduce_add_sync(full, st.nodes), leaves = __reduce_add_sync(full, st.leaves), culled = __reduce_add_sync(full, st.culled);
    const unsigned exact = __reduce_add_sync(full, st.nexact ? 1u : 0u), ties = __reduce_add_sync(full, tie ? 1u : 0u);
    const unsigned nears = __reduce_add_sync(full, near ? 1u : 0u), ovf = __reduce_or_sync(full, st.overflow ? 1u : 0u);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&tr.counters[3], (unsigned long long)nodes);
        atomicAdd(&tr.counters[4], (unsigned long long)leaves);
        if (culled) atomicAdd(&tr.counters[6], (unsigned long long)culled);
        if (exact) atomicAdd(&tr.counters[0], (unsigned long long)exact);
        if (ties) atomicAdd(&tr.counters[1], (unsigned long long)ties);
        if (nears) atomicAdd(&tr.counters[5], (unsigned long long)nears);
        if (ovf) atomicOr(&tr.counters[2], 1ull);
    }
}

// (8 blocks per SM = 64 registers: measured best of 3 / 4 / 5 / 6 / 8 / 10 on the config 3 and config 4 segment batches)
// out8: kAnyHit; out32: kCount, kAnyId; kList: ids of query i go to list[offsets[i] ..), and a count that disagrees with
// offsets[i+1] - offsets[i] (the caller's prefix sums of a kCount run over the same batch) raises counters[2] bit 2
template <int KIND, int MODE>
__global__ void __launch_bounds__(128, 8) intersect_kernel(const TreeRef tr, const double* __restrict__ q, uint64_t n, uint8_t* __restrict__ out8,
                                                           uint32_t* __restrict__ out32, const uint64_t* __restrict__ offsets, uint32_t* __restrict__ list) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    if (slot < n) {
        const uint64_t i = tr.perm ? tr.perm[slot] : slot;
        if (MODE == kList) {
            const uint64_t lo = offsets[i], hi = offsets[i + 1];
            uint32_t r = 0;
            // (a query with no room for any id still traverses: its count must be 0)
            if (hi > lo) r = traverse_intersect<KIND, kList>(tr.nodes, tr.tris, tr.root, ld3(q + 6 * i), ld3(q + 6 * i + 3), st, list + lo, (uint32_t)(hi - lo));
            else r = traverse_intersect<KIND, kAnyHit>(tr.nodes, tr.tris, tr.root, ld3(q + 6 * i), ld3(q + 6 * i + 3), st);
            if (r != hi - lo) atomicOr(&tr.counters[2], 4ull);
        } else {
            const uint32_t r = traverse_intersect<KIND, MODE>(tr.nodes, tr.tris, tr.root, ld3(q + 6 * i), ld3(q + 6 * i + 3), st);
            if (MODE == kAnyHit) out8[i] = (uint8_t)r; else out32[i] = r;
        }
    }
    flush_stats(tr, st, 0, 0);
}

// do_intersect / number_of_intersected_primitives / any_intersected_primitive by PERSISTENT threads with query replacement: the traversal of
// bvh.cuh::traverse_line (same node tests, same leaf routines, same results) where a lane takes the next query of the coherence order as soon as its own is
// answered.  On the config 3 shadow segments the thread-per-query kernel runs with 11 of 32 lanes active per instruction (profiles/r2_18): any-hit queries end
// at their first hit, long segments cross many nodes.  B200AABB_OPT_RAY_KERNEL selects it (modes any hit, count, any id; the list mode keeps the other kernel).
#ifndef B2_IS_BLOCKS
#define B2_IS_BLOCKS 8
#endif
template <int KIND, int MODE>
__global__ void __launch_bounds__(128, B2_IS_BLOCKS) intersect_stream_kernel(const TreeRef tr, const double* __restrict__ q, uint64_t n, unsigned long long* __restrict__ cursor,
                                                                  uint8_t* __restrict__ out8, uint32_t* __restrict__ out32) {
    const unsigned full = 0xFFFFFFFFu, lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    constexpr bool kStop = MODE == kAnyHit || MODE == kAnyId;
    constexpr int kPendCount = 6, kPendGoAt = kPendCount - 1;
    const float tlimit = KIND == kSegmentQ ? 1.0f : INFINITY, tstart = KIND == kLineQ ? -INFINITY : 0.0f;
    TravStat st;
    unsigned n_exact = 0;
    bool have = false, more = true;
    uint64_t qi = 0;
    Vec3 s = {0.0, 0.0, 0.0}, e = {0.0, 0.0, 0.0};
    LineCtx c;
    uint32_t stack[kStackCap];
    uint32_t pend[kPendCount + 1];
    int sp = 0, np = 0, nex_q = 0;
    uint32_t cur = kNone, found = 0, found_id = kNone;
    while (true) {
        const unsigned need = __ballot_sync(full, !have);
        if (more && need) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(need));
            base = __shfl_sync(full, base, leader);
            if (!have) {
                const unsigned long long pos = base + __popc(need & lt);
                if (pos < n) {
                    qi = tr.perm ? tr.perm[pos] : pos;
                    s = ld3(q + 6 * qi); e = ld3(q + 6 * qi + 3);
                    c = KIND == kLineQ ? make_line_pv(s, e) : make_line(s, e);
                    sp = 0; np = 0; nex_q = 0; cur = tr.root; found = 0; found_id = kNone;
                    have = true;
                }
            }
            if (base + __popc(need) >= n) more = false;
        }
        if (!__any_sync(full, have)) break;
        while (have && cur != kNone && np < kPendGoAt) {                  // node phase (bvh.cuh::traverse_line)
            const NodeView nd = load_node(tr.nodes, cur);
            ++st.nodes;
            float tn;
            const bool h0 = line_box(nd.lo0, nd.hi0, c, tlimit, tn, tstart);
            const bool h1 = line_box(nd.lo1, nd.hi1, c, tlimit, tn, tstart);
            uint32_t next = kNone;
            if (h0) { if (nd.c0 & kLeafBit) pend[np++] = nd.c0; else next = nd.c0; }
            if (h1) {
                if (nd.c1 & kLeafBit) pend[np++] = nd.c1;
                else if (next == kNone) next = nd.c1;
                else if (sp < kStackCap) stack[sp++] = nd.c1;
                else st.overflow = 1;
            }
            if (next == kNone && sp > 0) next = stack[--sp];
            cur = next;
        }
        if (have) {
            bool done = false;
            while (np > 0) {                                              // leaf phase
                uint32_t id;
                const Tri3 t = load_tri(tr.tris, pend[--np] & ~kLeafBit, id);
                ++st.leaves;
                const bool hit = KIND == kRayQ ? ray_tri_do_intersect(t, s, e, nex_q) : KIND == kLineQ ? line_tri_do_intersect(t, s, e, nex_q) : seg_tri_do_intersect(t, s, e, nex_q);
                if (hit) { ++found; found_id = id; if (kStop) { done = true; break; } }
            }
            if (done || cur == kNone) {
                if (MODE == kAnyHit) out8[qi] = (uint8_t)(found ? 1u : 0u);
                else out32[qi] = MODE == kAnyId ? found_id : found;
                n_exact += nex_q ? 1u : 0u;
                have = false;
            }
        }
    }
    const unsigned nodes = __reduce_add_sync(full, st.nodes), leaves = __reduce_add_sync(full, st.leaves);
    const unsigned exact = __reduce_add_sync(full, n_exact), ovf = __reduce_or_sync(full, st.overflow ? 1u : 0u);
    if (lane == 0) {
        atomicAdd(&tr.counters[3], (unsigned long long)nodes);
        atomicAdd(&tr.counters[4], (unsigned long long)leaves);
        if (exact) atomicAdd(&tr.counters[0], (unsigned long long)exact);
        if (ovf) atomicOr(&tr.counters[2], 1ull);
    }
}

// Side_of_triangle_mesh (Side_of_triangle_mesh.h:265-306): -1 ON_UNBOUNDED_SIDE, 0 ON_BOUNDARY, +1 ON_BOUNDED_SIDE of a closed mesh.  A point outside
// the tree's box is outside; otherwise rays are cast until one is decisive: first the vertical one towards the nearer z face of the box
// (Point_inside_vertical_ray_cast.h:79-81), then fixed pseudo-random directions — the same for every query, as the reference's re-seeded generator
// gives (:85-91; its particular sequence does not matter: the answer does not depend on the ray, only the number of attempts does).
// out_attempts (optional) = rays cast per query.
__device__ __forceinline__ Vec3 side_direction(int attempt, double pz, const double* box) {
    if (attempt == 0) return {0.0, 0.0, 2.0 * pz < box[5] + box[2] ? -1.0 : 1.0};
    unsigned long long z = 0x9E3779B97F4A7C15ull * (unsigned long long)attempt;       // SplitMix64 of the attempt number -> a direction in [-1, 1]^3
    double c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        z += 0x9E3779B97F4A7C15ull;
        unsigned long long x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
        c[k] = (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    }
    if (fabs(c[0]) + fabs(c[1]) + fabs(c[2]) < 0.25) c[attempt % 3] = 1.0;
    return {c[0], c[1], c[2]};
}
constexpr int kSideMaxAttempts = 64;
// XF = AffXf: the mesh stands under the affine transformation x.xf (the points are given in the frame of the transformed mesh); x.box = a box that contains
// the image of the mesh, x.mag = the largest coordinate magnitude of that box
struct SideXf { AffXf xf; double box[6]; double mag; };
template <class XF>
__global__ void __launch_bounds__(128, 8) side_of_mesh_kernel(const TreeRef tr, const double* __restrict__ pts, uint64_t n, const double* __restrict__ rootbox,
                                                              int8_t* __restrict__ out, uint8_t* __restrict__ out_attempts, const SideXf x) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    if (slot < n) {
        const uint64_t i = tr.perm ? tr.perm[slot] : slot;
        const Vec3 p = ld3(pts + 3 * i);
        double box[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) box[k] = XF::kIdentity ? rootbox[k] : x.box[k];
        int side = -1, attempts = 0;
        if (!(p.x < box[0] || p.x > box[3] || p.y < box[1] || p.y > box[4] || p.z < box[2] || p.z > box[5])) {
            const double scale = fmax(1.0, XF::kIdentity ? tr.mag : x.mag);          // p + d must differ from p
            int r = kSideIndeterminate;
            for (; attempts < kSideMaxAttempts && r == kSideIndeterminate; ++attempts) {
                const Vec3 d = side_direction(attempts, p.z, box);
                const Vec3 q = {p.x + d.x * scale, p.y + d.y * scale, p.z + d.z * scale};
                if constexpr (XF::kIdentity) r = traverse_side(tr.nodes, tr.tris, tr.root, p, q, st);
                else r = traverse_side<AffXf>(tr.nodes, tr.tris, tr.root, p, q, st, x.xf);
            }
            if (r == kSideIndeterminate) st.overflow = 1;   // (64 degenerate rays in a row: reported, never guessed)
            side = r == kSideBoundary ? 0 : ((r & 1) ? 1 : -1);
        }
        out[i] = (int8_t)side;
        if (out_attempts) out_attempts[i] = (uint8_t)attempts;
    }
    flush_stats(tr, st, 0, 0);
}

// the same for Plane_3 queries: q = n x 4 coefficients (a, b, c, d)
template <int MODE>
__global__ void __launch_bounds__(128, 8) plane_kernel(const TreeRef tr, const double* __restrict__ q, uint64_t n, uint8_t* __restrict__ out8, uint32_t* __restrict__ out32,
                                                       const uint64_t* __restrict__ offsets, uint32_t* __restrict__ list) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    if (i < n) {
        const double a = __ldg(q + 4 * i), b = __ldg(q + 4 * i + 1), c = __ldg(q + 4 * i + 2), d = __ldg(q + 4 * i + 3);
        if (MODE == kList) {
            const uint64_t lo = offsets[i], hi = offsets[i + 1];
            uint32_t r = 0;
            if (hi > lo) r = traverse_plane<kList>(tr.nodes, tr.tris, tr.root, a, b, c, d, st, list + lo, (uint32_t)(hi - lo));
            else r = traverse_plane<kAnyHit>(tr.nodes, tr.tris, tr.root, a, b, c, d, st);
            if (r != hi - lo) atomicOr(&tr.counters[2], 4ull);
        } else {
            const uint32_t r = traverse_plane<MODE>(tr.nodes, tr.tris, tr.root, a, b, c, d, st);
            if (MODE == kAnyHit) out8[i] = (uint8_t)r; else out32[i] = r;
        }
    }
    flush_stats(tr, st, 0, 0);
}

// the same for Triangle_3 queries: query i = the 9 doubles at q + stride * i (stride 9: the caller's batch; 12: the TriRec array of another tree)
template <int MODE, bool SEG>
__global__ void __launch_bounds__(128, 6) triangle_kernel(const TreeRef tr, const double* __restrict__ q, uint32_t stride, uint64_t n, uint8_t* __restrict__ out8,
                                                          uint32_t* __restrict__ out32, const uint64_t* __restrict__ offsets, uint32_t* __restrict__ list) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    if (slot < n) {
        const uint64_t i = tr.perm ? tr.perm[slot] : slot;
        const double* p = q + (uint64_t)stride * i;
        const Tri3 t = {ld3(p), ld3(p + 3), ld3(p + 6)};
        if (MODE == kList) {
            const uint64_t lo = offsets[i], hi = offsets[i + 1];
            uint32_t r = 0;
            if (hi > lo) r = traverse_triangle<kList, NoXf, SEG>(tr.nodes, tr.tris, tr.root, t, st, list + lo, (uint32_t)(hi - lo));
            else r = traverse_triangle<kAnyHit, NoXf, SEG>(tr.nodes, tr.tris, tr.root, t, st);
            if (r != hi - lo) atomicOr(&tr.counters[2], 4ull);
        } else {
            const uint32_t r = traverse_triangle<MODE, NoXf, SEG>(tr.nodes, tr.tris, tr.root, t, st);
            if (MODE == kAnyHit) out8[i] = (uint8_t)r; else out32[i] = r;
        }
    }
    flush_stats(tr, st, 0, 0);
}

// tree.do_intersect(other_tree) (AABB_traits_3.h:679-690: every primitive of the traversed tree is a Triangle_3 query against the other tree, the
// traversal stops at the first pair): one thread per primitive of `a` (its Morton order is the coherence order), traversing `tr` = the other tree;
// witness[0] = 1 at the first intersecting pair found, witness[1] / [2] = ids of one such pair (which one is unspecified).  Threads that start after a
// pair has been found return at once.
// XF: both trees stand under affine transformations (Rigid_triangle_mesh_collision_detection::does_A_intersect_B, :149-161): x.q transforms the primitives
// that serve as queries, x.b the traversed tree (its primitives at the leaves, its boxes at the nodes, bvh.cuh::AffXf).
struct XfPair { AffXf q, b; };
template <bool XF>
__global__ void __launch_bounds__(128, 6) tree_tree_kernel(const TreeRef tr, const TriRec* __restrict__ a, uint64_t na, unsigned* __restrict__ witness, const XfPair x) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    if (i < na && *(volatile unsigned*)witness == 0u) {
        uint32_t ida;
        Tri3 t = load_tri(a, (uint32_t)i, ida);
        uint32_t idb;
        if (XF) { t = x.q.apply(t); idb = traverse_triangle<kAnyId, AffXf>(tr.nodes, tr.tris, tr.root, t, st, nullptr, 0, x.b); }
        else idb = traverse_triangle<kAnyId>(tr.nodes, tr.tris, tr.root, t, st);
        if (idb != kNone && atomicExch(&witness[0], 1u) == 0u) { witness[1] = ida; witness[2] = idb; }
    }
    flush_stats(tr, st, 0, 0);
}

// all_intersections / any_intersection (Ray_3): bvh.cuh::traverse_ray_objects.  kCount: out32[i]; kList: entries of ray i at offsets[i] ..
// (a count that disagrees with offsets[i+1] - offsets[i] raises counters[2] bit 2); kAnyId: prim / flags / pt / pt2 like first_intersection
constexpr int kPlaneObjects = 3;        // KIND of ray_objects_kernel for Plane_3 queries (next to LineKind's 0..2)
template <int MODE, int KIND>
__global__ void __launch_bounds__(128, 5) ray_objects_kernel(const TreeRef tr, const double* __restrict__ rays, uint64_t n, uint32_t* __restrict__ out32,
                                                             const uint64_t* __restrict__ offsets, uint32_t* __restrict__ ids, uint8_t* __restrict__ kinds,
                                                             double* __restrict__ pts, uint32_t* __restrict__ prim, uint8_t* __restrict__ flags,
                                                             double* __restrict__ pt, double* __restrict__ pt2) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    if (slot < n) {
        const uint64_t i = tr.perm ? tr.perm[slot] : slot;
        constexpr bool kPlane = KIND == kPlaneObjects;        // Plane_3: four coefficients per query, its own traversal
        const Vec3 s = kPlane ? ld3(rays + 4 * i) : ld3(rays + 6 * i), q = kPlane ? Vec3{rays[4 * i + 3], 0.0, 0.0} : ld3(rays + 6 * i + 3);
        auto objects = [&](auto mode, uint32_t* oi, uint8_t* ok, double* op, uint32_t cap) -> uint32_t {
            constexpr int M = decltype(mode)::value;
            if constexpr (kPlane) return traverse_plane_objects<M>(tr.nodes, tr.tris, tr.root, s.x, s.y, s.z, q.x, st, oi, ok, op, cap);
            else return traverse_ray_objects<M, kPlane ? kRayQ : KIND>(tr.nodes, tr.tris, tr.root, s, q, st, oi, ok, op, cap);
        };
        if (MODE == kCount) {
            out32[i] = objects(std::integral_constant<int, kCount>(), nullptr, nullptr, nullptr, 0);
        } else if (MODE == kList) {
            const uint64_t lo = offsets[i], hi = offsets[i + 1];
            const uint32_t r = objects(std::integral_constant<int, kList>(), ids + lo, kinds + lo, pts + 6 * lo, (uint32_t)(hi - lo));
            if (r != hi - lo) atomicOr(&tr.counters[2], 4ull);
        } else {
            uint32_t id = kNone; uint8_t kind = 0; double p[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            objects(std::integral_constant<int, kAnyId>(), &id, &kind, p, 1);
            if (prim) prim[i] = id;
            if (flags) flags[i] = (uint8_t)((kind ? B200AABB_FLAG_HIT : 0) | (kind == 2 ? B200AABB_FLAG_SEGMENT : 0) | (kind == 3 ? B200AABB_FLAG_TRIANGLE : 0) | (st.nexact ? B200AABB_FLAG_EXACT : 0));
            if (pt) { pt[3 * i] = p[0]; pt[3 * i + 1] = p[1]; pt[3 * i + 2] = p[2]; }
            if (pt2) { pt2[3 * i] = p[3]; pt2[3 * i + 1] = p[4]; pt2[3 * i + 2] = p[5]; }
        }
    }
    flush_stats(tr, st, 0, 0);
}

// The same traversal as bvh.cuh::traverse_first (operation for operation: the results are bit-identical) run by PERSISTENT threads that take the next ray of
// the coherence order from a cursor as soon as their ray is done ("while-while with replacement", Aila & Laine 2009): in the thread-per-ray kernel a warp
// runs until its slowest ray ends, with 17 of 32 lanes active per instruction on the benchmark batch (profiles/r2_08) — half the pinhole rays miss the terrain.
// B200AABB_OPT_RAY_KERNEL selects it; rays with a skip list and one-primitive trees use the thread-per-ray kernel.
__global__ void __launch_bounds__(128, 5) first_intersection_stream_kernel(const TreeRef tr, const double* __restrict__ rays, uint64_t n, unsigned long long* __restrict__ cursor,
                                                                         uint32_t* __restrict__ prim, double* __restrict__ tpar, double* __restrict__ pt, double* __restrict__ pt2,
                                                                         uint8_t* __restrict__ flags) {
    const unsigned full = 0xFFFFFFFFu, lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    TravStat st;
    unsigned n_exact = 0, n_tie = 0, n_near = 0;
    bool have = false, more = true;
    uint64_t qi = 0;
    Vec3 s = {0.0, 0.0, 0.0}, q = {0.0, 0.0, 0.0};
    LineCtx c;
    FirstHit best;
    double best_u = 0.0, strict_lo = DBL_MAX, tie_lo = DBL_MAX;
    bool tie = false;
    float tlimit = INFINITY;
    uint32_t stack[kStackCap];
    float stack_t[kStackCap];
    int sp = 0, np = 0, nex_q = 0;
    uint32_t pend[2];
    float pend_t[2];
    uint32_t cur = kNone;
    float cur_t = 0.0f;
    while (true) {
        const unsigned need = __ballot_sync(full, !have);
        if (more && need) {                                               // idle lanes take the next rays of the order
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(need));
            base = __shfl_sync(full, base, leader);
            if (!have) {
                const unsigned long long pos = base + __popc(need & lt);
                if (pos < n) {
                    qi = tr.perm ? tr.perm[pos] : pos;
                    s = ld3(rays + 6 * qi); q = ld3(rays + 6 * qi + 3);
                    c = make_line(s, q, tr.mag);
                    best.prim = kNone; best.kind = 0; best.t = DBL_MAX; best.p0 = s; best.p1 = s; best.tie = 0; best.near = 0;
                    best_u = 0.0; strict_lo = DBL_MAX; tie_lo = DBL_MAX; tie = false; tlimit = INFINITY;
                    sp = 0; np = 0; nex_q = 0; cur = tr.root; cur_t = 0.0f;
                    have = true;
                }
            }
            if (base + __popc(need) >= n) more = false;
        }
        if (!__any_sync(full, have)) break;
        while (have && cur != kNone && np == 0) {                         // node phase (bvh.cuh::traverse_first)
            const NodeView nd = load_node(tr.nodes, cur);
            ++st.nodes;
            float t0, t1;
            bool h0 = line_box(nd.lo0, nd.hi0, c, tlimit, t0);
            bool h1 = line_box(nd.lo1, nd.hi1, c, tlimit, t1);
            uint32_t c0 = nd.c0, c1 = nd.c1;
            if (h0 && h1 && t1 < t0) { const uint32_t tc = c0; c0 = c1; c1 = tc; const float tt = t0; t0 = t1; t1 = tt; }
            else if (!h0 && h1) { c0 = c1; t0 = t1; h0 = true; h1 = false; }
            uint32_t next = kNone;
            if (h0) { if (c0 & kLeafBit) { pend[np] = c0; pend_t[np] = t0; ++np; } else { next = c0; cur_t = t0; } }
            if (h1) {
                if (c1 & kLeafBit) { pend[np] = c1; pend_t[np] = t1; ++np; }
                else if (next == kNone) { next = c1; cur_t = t1; }
                else if (sp < kStackCap) { stack[sp] = c1; stack_t[sp] = t1; ++sp; }
                else st.overflow = 1;
            }
            while (next == kNone && sp > 0) {
                --sp;
                if (stack_t[sp] <= tlimit) { next = stack[sp]; cur_t = stack_t[sp]; }
            }
            cur = next;
        }
        if (have) {
            for (int k = 0; k < np; ++k) {                                // leaf phase
                if (!(pend_t[k] <= tlimit)) continue;
                uint32_t id;
                const Tri3 t = load_tri(tr.tris, pend[k] & ~kLeafBit, id);
                ++st.leaves;
                int nex = 0;
                const RayX x = ray_tri_intersection_call(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, s.x, s.y, s.z, q.x, q.y, q.z, c.major, &nex);
                nex_q += nex;
                if (x.kind == 0) continue;
                if (x.t < best.t || (x.t == best.t && id < best.prim)) {
                    if (best.prim != kNone) {
                        if (x.t < best.t) { strict_lo = fmin(strict_lo, fmin(best.t - best_u, tie_lo)); tie = false; tie_lo = DBL_MAX; }
                        else { tie = true; tie_lo = fmin(tie_lo, best.t - best_u); }
                    }
                    best.prim = id; best.kind = x.kind; best.t = x.t; best.p0 = x.p0; best.p1 = x.p1; best_u = x.u;
                    tlimit = float_above(x.t + ((fabs(x.t) * 2e-6 + x.u) + c.t_abs_slack));
                } else if (x.t == best.t) { tie = true; tie_lo = fmin(tie_lo, x.t - x.u); }
                else strict_lo = fmin(strict_lo, x.t - x.u);
            }
            np = 0;
            if (cur != kNone && cur_t > tlimit) {
                cur = kNone;
                while (cur == kNone && sp > 0) {
                    --sp;
                    if (stack_t[sp] <= tlimit) { cur = stack[sp]; cur_t = stack_t[sp]; }
                }
            }
            if (cur == kNone) {                                           // this ray is done: results out, the lane is free
                const bool hit = best.prim != kNone;
                const bool is_tie = hit && tie, is_near = hit && strict_lo <= best.t + best_u;
                if (prim) prim[qi] = best.prim;
                if (tpar) tpar[qi] = hit ? best.t : 0.0;
                if (pt) { pt[3 * qi] = hit ? best.p0.x : 0.0; pt[3 * qi + 1] = hit ? best.p0.y : 0.0; pt[3 * qi + 2] = hit ? best.p0.z : 0.0; }
                if (pt2) { const bool seg = hit && best.kind == 2; pt2[3 * qi] = seg ? best.p1.x : 0.0; pt2[3 * qi + 1] = seg ? best.p1.y : 0.0; pt2[3 * qi + 2] = seg ? best.p1.z : 0.0; }
                if (flags) flags[qi] = (uint8_t)((hit ? B200AABB_FLAG_HIT : 0) | (hit && best.kind == 2 ? B200AABB_FLAG_SEGMENT : 0) | (is_tie ? B200AABB_FLAG_TIE : 0) |
                                                 (nex_q ? B200AABB_FLAG_EXACT : 0) | (is_near ? B200AABB_FLAG_NEAR_TIE : 0));
                n_exact += nex_q ? 1u : 0u; n_tie += is_tie ? 1u : 0u; n_near += is_near ? 1u : 0u;
                have = false;
            }
        }
    }
    const unsigned nodes = __reduce_add_sync(full, st.nodes), leaves = __reduce_add_sync(full, st.leaves);
    const unsigned exact = __reduce_add_sync(full, n_exact), ties = __reduce_add_sync(full, n_tie), nears = __reduce_add_sync(full, n_near), ovf = __reduce_or_sync(full, st.overflow ? 1u : 0u);
    if (lane == 0) {
        atomicAdd(&tr.counters[3], (unsigned long long)nodes);
        atomicAdd(&tr.counters[4], (unsigned long long)leaves);
        if (exact) atomicAdd(&tr.counters[0], (unsigned long long)exact);
        if (ties) atomicAdd(&tr.counters[1], (unsigned long long)ties);
        if (nears) atomicAdd(&tr.counters[5], (unsigned long long)nears);
        if (ovf) atomicOr(&tr.counters[2], 1ull);
    }
}

template <bool SKIP>
__global__ void __launch_bounds__(128, 5) first_intersection_kernel(const TreeRef tr, const double* __restrict__ rays, uint64_t n, uint32_t* __restrict__ prim,
                                                                 double* __restrict__ tpar, double* __restrict__ pt, double* __restrict__ pt2,
                                                                 uint8_t* __restrict__ flags, const uint32_t* __restrict__ skip_ids,
                                                                 const uint32_t* __restrict__ skip_mask, uint32_t mask_words) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    int tie = 0, near = 0;
    if (slot < n) {
        const uint64_t i = tr.perm ? tr.perm[slot] : slot;
        const FirstHit h = SKIP ? traverse_first<true>(tr.nodes, tr.tris, tr.root, ld3(rays + 6 * i), ld3(rays + 6 * i + 3), tr.mag, st, skip_ids ? skip_ids[i] : kNone, skip_mask, mask_words)
                                : traverse_first<false>(tr.nodes, tr.tris, tr.root, ld3(rays + 6 * i), ld3(rays + 6 * i + 3), tr.mag, st);
        const bool hit = h.prim != kNone;
        tie = h.tie; near = h.near;
        if (prim) prim[i] = h.prim;
        if (tpar) tpar[i] = hit ? h.t : 0.0;
        if (pt) { pt[3 * i] = hit ? h.p0.x : 0.0; pt[3 * i + 1] = hit ? h.p0.y : 0.0; pt[3 * i + 2] = hit ? h.p0.z : 0.0; }
        if (pt2) { const bool seg = hit && h.kind == 2; pt2[3 * i] = seg ? h.p1.x : 0.0; pt2[3 * i + 1] = seg ? h.p1.y : 0.0; pt2[3 * i + 2] = seg ? h.p1.z : 0.0; }
        if (flags) flags[i] = (uint8_t)((hit ? B200AABB_FLAG_HIT : 0) | (hit && h.kind == 2 ? B200AABB_FLAG_SEGMENT : 0) | (h.tie ? B200AABB_FLAG_TIE : 0) |
                                        (st.nexact ? B200AABB_FLAG_EXACT : 0) | (h.near ? B200AABB_FLAG_NEAR_TIE : 0));
    }
    flush_stats(tr, st, tie, near);
}

template <bool SEG>
__global__ void __launch_bounds__(128) closest_point_kernel(const TreeRef tr, const double* __restrict__ pts, uint64_t n, double* __restrict__ cp,
                                                            uint32_t* __restrict__ prim, double* __restrict__ sqd, uint8_t* __restrict__ flags) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TravStat st;
    int tie = 0, near = 0;
    if (slot < n) {
        const uint64_t i = tr.perm ? tr.perm[slot] : slot;
        const Closest c = traverse_closest<SEG>(tr.nodes, tr.tris, tr.root, ld3(pts + 3 * i), tr.mag, st);
        tie = c.tie; near = c.near;
        if (cp) { cp[3 * i] = c.p.x; cp[3 * i + 1] = c.p.y; cp[3 * i + 2] = c.p.z; }
        if (prim) prim[i] = c.prim;
        if (sqd) sqd[i] = c.sqd;
        if (flags) flags[i] = (uint8_t)((c.tie ? B200AABB_FLAG_TIE : 0) | (st.nexact ? B200AABB_FLAG_EXACT : 0) | (c.near ? B200AABB_FLAG_NEAR_TIE : 0));
    }
    flush_stats(tr, st, tie, near);
}

// ---- predicate self-test hook (b200aabb_debug_predicate): evaluates ONE device predicate per record so that the -m gpu
// tests can compare the device-compiled filters and exact stage with the oracle on adversarial inputs
__global__ void __launch_bounds__(128) predicate_kernel(int which, const double* __restrict__ in, int stride, uint64_t n, int* __restrict__ out,
                                                        int* __restrict__ used_exact, double* __restrict__ dout) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = in + (uint64_t)stride * i;
    int nexact = 0, r = 0;
    auto V = [&](int k) { return Vec3{p[3 * k], p[3 * k + 1], p[3 * k + 2]}; };
    switch (which) {
        case 0: r = orient3(V(0), V(1), V(2), V(3), nexact); break;
        case 1: r = orient3_dir(V(0), V(1), V(2), V(3), V(4), nexact); break;
        case 2: r = coplanar_orient(V(0), V(1), V(2), nexact); break;
        case 3: r = cmp_dist(V(0), V(1), V(2), nexact); break;
        case 4: r = seg_tri_do_intersect(Tri3{V(0), V(1), V(2)}, V(3), V(4), nexact); break;
        case 5: r = ray_tri_do_intersect(Tri3{V(0), V(1), V(2)}, V(3), V(4), nexact); break;
        case 6: {
            const Vec3 s = V(3), q = V(4), d = q - s;
            const RayX x = ray_tri_intersection(Tri3{V(0), V(1), V(2)}, s, q, d, ray_major_axis(d), nexact);
            r = x.kind;
            if (dout) { double* o = dout + 8 * i; o[0] = x.t; o[1] = x.p0.x; o[2] = x.p0.y; o[3] = x.p0.z; o[4] = x.p1.x; o[5] = x.p1.y; o[6] = x.p1.z; o[7] = x.u; }
            break;
        }
        case 7: {
            const Vec3 c = closest_on_triangle(Tri3{V(0), V(1), V(2)}, V(3));
            if (dout) { double* o = dout + 8 * i; o[0] = c.x; o[1] = c.y; o[2] = c.z; }
            break;
        }
        case 8: r = exact_det3_sign(dd_diff(p[3], p[0]), dd_diff(p[4], p[1]), dd_diff(p[5], p[2]), dd_diff(p[6], p[0]), dd_diff(p[7], p[1]), dd_diff(p[8], p[2]),
                                    dd_diff(p[9], p[0]), dd_diff(p[10], p[1]), dd_diff(p[11], p[2])); break;
        default: break;
    }
    out[i] = r;
    if (used_exact) used_exact[i] = nexact;
}

// max and arg max of n non-negative doubles (squared distances): non-negative binary64 values order like their bit patterns, so one 64-bit
// atomicMax per block does it; the index rides along in a second pass over the (few) maxima — ties -> smallest index
__global__ void __launch_bounds__(256) max_sqdist_kernel(const double* __restrict__ v, uint64_t n, unsigned long long* __restrict__ best_bits) {
    unsigned long long m = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = max(m, (unsigned long long)__double_as_longlong(v[i]));
    for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(best_bits, m);
}
__global__ void __launch_bounds__(256) argmax_sqdist_kernel(const double* __restrict__ v, uint64_t n, const unsigned long long* __restrict__ best_bits,
                                                            unsigned long long* __restrict__ best_index) {
    const unsigned long long b = *best_bits;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if ((unsigned long long)__double_as_longlong(v[i]) == b) atomicMin(best_index, (unsigned long long)i);
}

// ---- query coherence keys (sorted with radix_sort.cuh; the traversal kernels then follow the permutation) ------------
// Point_3: 30-bit Morton code (10 bits per axis) of the point in the isotropic frame of the root box
__global__ void __launch_bounds__(256) point_keys_kernel(const double* __restrict__ pts, uint32_t stride, uint64_t n, const double* __restrict__ rootbox,
                                                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lo[3], scale[3];
    morton_frame(rootbox[0], rootbox[1], rootbox[2], rootbox[3], rootbox[4], rootbox[5], lo, scale);
    const double* p = pts + (uint64_t)stride * i;                      // (stride 9: Triangle_3 queries, keyed by their first vertex)
    keys[i] = morton63(p[0], p[1], p[2], lo, scale) >> 33;
    vals[i] = (uint32_t)i;
}
// Segment_3 / Ray_3: 48-bit key interleaving 8 bits of each origin coordinate (root-box frame) and of each component of
// the max-norm-normalised direction, most significant bits first
__global__ void __launch_bounds__(256) line_keys_kernel(const double* __restrict__ q, uint64_t n, const double* __restrict__ rootbox,
                                                        uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int second_is_vector) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lo[3], scale[3];
    morton_frame(rootbox[0], rootbox[1], rootbox[2], rootbox[3], rootbox[4], rootbox[5], lo, scale);
    const double* r = q + 6 * i;
    const double d[3] = {second_is_vector ? r[3] : r[3] - r[0], second_is_vector ? r[4] : r[4] - r[1], second_is_vector ? r[5] : r[5] - r[2]};
    const double dmax = fmax(fabs(d[0]), fmax(fabs(d[1]), fabs(d[2])));
    uint32_t c[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double f = (r[k] - lo[k]) * scale[k] * (256.0 / 2097152.0);          // root box -> [0, 256)
        c[k] = (uint32_t)fmin(fmax(f, 0.0), 255.0);
        const double g = dmax > 0.0 ? (d[k] / dmax + 1.0) * 127.5 : 0.0;             // [-1, 1] -> [0, 255]
        c[3 + k] = (uint32_t)fmin(fmax(g, 0.0), 255.0);
    }
    uint64_t key = 0;
#pragma unroll
    for (int b = 7; b >= 0; --b)
#pragma unroll
        for (int k = 0; k < 6; ++k) key = (key << 1) | ((c[k] >> b) & 1u);
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// Handle
// ------------------------------------------------------------------------------------------------------------------
// HOST-pointer batches flow through THREE streams — uploads, kernels, downloads — each in issue order, tied together per staging slot by
// events (upload done -> kernels; kernels done -> download; kernels done -> the slot's next upload; download done -> the slot's next
// kernels).  The kernels of a batch fill the GPU one after the other anyway, and both copy engines work through their queues behind them;
// with one stream per chunk (round 1) the chunks' kernels queued up behind a long closest_point launch and — streams sharing hardware
// queues — held back uploads that had nothing to wait for (profiles/r2_05).
constexpr int kLineSlots = 8;                   // staging slots of ray / segment chunks: a whole batch can be uploaded while a closest_point launch runs
constexpr int kPointSlots = 2;                  // staging slots of point batches (a point batch stays whole, see run_batch)
constexpr int kPipeSlots = kLineSlots + kPointSlots;
constexpr int kHostScratch = 0, kDeviceScratch = 1, kScratchSlots = 2;   // sort / work-pool scratch: HOST batches (their kernels share one stream), DEVICE batches

struct b200aabb_tree {
    // Every entry point that touches the handle takes this lock: concurrent calls on ONE handle (the reference's callers run
    // tbb::parallel_for over const queries, AABB_tree.h:593-612) are serialised, never interleaved on the shared staging / scratch buffers.
    std::recursive_mutex mu;
    bool imported = false;         // the tree came from b200aabb_import_alloc: there are no primitives to rebuild from
    bool segments = false;         // the primitives are Segment_3 (b200aabb_set_segments), stored as the records (source, target, target)
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // primitives as inserted (device): 9 doubles each, optional ids
    double* d_soup = nullptr;
    uint32_t* d_ids = nullptr;
    uint64_t n = 0;
    // built tree
    bool built = false;
    BvhNode* d_nodes = nullptr;
    TriRec* d_tris = nullptr;
    TriF* d_trisf = nullptr;       // binary32 copies (pre-tests of the warp-cooperative closest_point kernel)
    double* d_rootbox = nullptr;   // 6 doubles
    uint32_t* d_leaf_parent = nullptr;   // parent node of every leaf (refit)
    unsigned long long* d_bounds = nullptr;   // scene bounds as ordered integers (build.cuh); valid when bounds_valid
    bool bounds_valid = false;     // b200aabb_set_mesh computed them while gathering the primitives
    int opt_exact_keys = 0;        // 1: always sort full 63-bit keys (the build's fallback path; test hook, option 102)
    int last_build_packed = 0; unsigned long long last_build_dups = 0;
    // Buffers of the build kept between builds (grow only; released by clear / destroy): cudaMalloc + cudaFree of a few GB cost 10 - 300 ms per
    // build on this driver, five to a hundred times the kernels (profiles/README.md, r2_12).  tree_cap = primitives the tree buffers have room for.
    struct BuildScratch { uint64_t* keys[2] = {nullptr, nullptr}; uint32_t* vals[2] = {nullptr, nullptr}; void* ws = nullptr; RefitEntry* frontier = nullptr;
                          unsigned* frontier_count = nullptr; uint32_t* rest = nullptr; unsigned long long* flags = nullptr; uint64_t cap = 0; uint64_t vals_cap = 0; } bs;
    uint64_t tree_cap = 0;
    int opt_ray_kernel = 0;        // first_intersection: 0 = one thread per ray, 1 = persistent threads with ray replacement (B200AABB_OPT_RAY_KERNEL)
    unsigned long long* d_ray_cursor = nullptr;   // [kScratchSlots]
    int ray_grid = 0, seg_grid = 0;
    int opt_classic_build = 0;     // 1: hierarchy_kernel for every node + refit_local_kernel instead of the fused bottom-up kernel (the fallback of a full frontier; test hook, option 103)
    int last_build_classic = 0;
    uint32_t root = kNone;
    double h_bbox[6];
    double mag = 0.0;              // largest coordinate magnitude of the root box
    unsigned long long* d_counters = nullptr;   // [8]
    unsigned* d_witness = nullptr;              // [3]: flag + ids of one intersecting pair of a tree-vs-tree test
    // options
    int opt_sort_queries = 1;      // B200AABB_OPT_SORT_QUERIES
    int opt_closest_kernel = 1;    // B200AABB_OPT_CLOSEST_KERNEL: 1 = warp-cooperative pool (closest_warp.cuh), 2 = lane pairs with private stacks (closest_pair.cuh; measured slower, profiles/r2_02), 0 = one thread per query
    int opt_warp_exact_go = 0;       // B200AABB_OPT_WARP_EXACT_GO (0 = the kernel's default: 64 pair kernel, 160 pool kernel)
    int opt_host_parts = 0;        // B200AABB_OPT_HOST_PARTS (0 = automatic)
    int opt_debug_refit_frontier = kRefitFrontierPerBlock;   // test hook (option 101): frontier entries per block of the refit (0: every hand-over overflows)
    // warp-cooperative closest_point: per pipeline slot a batch cursor and the pools' spill area; grid = SMs x resident warps
    struct WarpScratch { unsigned long long* cursor = nullptr; uint2* spill = nullptr; WarpSlots64* slots64 = nullptr;
                         uint32_t* fc_cand = nullptr; uint8_t* fc_ncand = nullptr; uint32_t* fc_overflow = nullptr; unsigned long long* fc_count = nullptr; uint64_t fc_cap = 0; } warp[kScratchSlots];   // fc_*: the two-pass kernels (closest_filter_kernel)
    int warp_grid = 0;
    int pair_grid = 0;             // grid of closest_point_pair_kernel (closest_pair.cuh); it shares the scratch of the pool kernel
    // query coherence sort scratch, one per pipeline slot (slot 0 also serves DEVICE-pointer batches); grow-only
    struct SortScratch { uint64_t* keys[2] = {nullptr, nullptr}; uint32_t* vals[2] = {nullptr, nullptr}; void* ws = nullptr; uint64_t cap = 0; } sort[kScratchSlots];
    // staging for HOST-pointer batches (two pipeline slots)
    void* d_stage_in[kPipeSlots] = {};
    void* d_stage_out[kPipeSlots] = {};
    size_t stage_in_bytes[kPipeSlots] = {}, stage_out_bytes[kPipeSlots] = {};
    cudaStream_t s_up = nullptr, s_run = nullptr, s_down = nullptr;   // uploads, kernels, downloads of HOST batches
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_up[kPipeSlots] = {}, ev_run[kPipeSlots] = {}, ev_down[kPipeSlots] = {};
    int next_line_slot = 0, next_point_slot = 0, cur_slot = 0;   // pipeline slot of the next chunk / of the chunk being launched
    uint64_t cur_offset = 0;           // first query of the chunk being launched (HOST batches)
    uint32_t* d_stage_aux[kPipeSlots] = {}; size_t stage_aux_bytes = 0;   // per-query side input of a HOST batch (skip ids)
    int opt_async_host = 0;        // B200AABB_OPT_ASYNC_HOST
    bool async_pending = false;
    struct TraceRec { cudaEvent_t e[4]; int slot; uint64_t cnt; int shape; };   // B200AABB_PIPE_TRACE=1: per-chunk events of the HOST pipeline (diagnostic)
    std::vector<TraceRec> trace;
    b200aabb_stats stats;
    bool query_timing_pending = false;
};

namespace {

#define B2_CUDA(t, call)                                                                                         \
    do {                                                                                                         \
        cudaError_t e_ = (call);                                                                                 \
        if (e_ != cudaSuccess) {                                                                                 \
            (t)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                       \
            return e_ == cudaErrorMemoryAllocation ? B200AABB_ERR_OUT_OF_MEMORY : B200AABB_ERR_CUDA;              \
        }                                                                                                        \
    } while (0)

constexpr uint64_t kChunkMax = 1ull << 22;   // most queries per pipeline slot for HOST-pointer batches of rays / segments
constexpr uint64_t kPointChunkMax = 1ull << 24;   // ... of points
constexpr uint64_t kChunkMin = 1ull << 18;
// a HOST batch is cut into about `opt_host_parts` chunks so that only the first chunk's upload and the last chunk's download are exposed
uint64_t chunk_for(uint64_t n, int parts, uint64_t cap) { const uint64_t p = (uint64_t)std::max(1, parts); return std::min<uint64_t>(n, std::min(cap, std::max(kChunkMin, ((n + p - 1) / p + 32767) / 32768 * 32768))); }

#define B2_LOCK(t) std::unique_lock<std::recursive_mutex> lock_; if (t) lock_ = std::unique_lock<std::recursive_mutex>(const_cast<b200aabb_tree*>(t)->mu)

int sync_pipeline(b200aabb_tree* t);
unsigned grid_for(uint64_t n, int block);

int fail(b200aabb_tree* t, int code, const char* msg) { if (t) t->err = msg; return code; }

void free_tree_buffers(b200aabb_tree* t) {
    cudaFree(t->d_nodes); cudaFree(t->d_tris); cudaFree(t->d_trisf); cudaFree(t->d_leaf_parent);
    t->d_nodes = nullptr; t->d_tris = nullptr; t->d_trisf = nullptr; t->d_leaf_parent = nullptr; t->built = false; t->root = kNone; t->tree_cap = 0;
}
void free_build_scratch(b200aabb_tree* t) {
    auto& b = t->bs;
    cudaFree(b.keys[0]); cudaFree(b.keys[1]); cudaFree(b.vals[0]); cudaFree(b.vals[1]); cudaFree(b.ws); cudaFree(b.frontier); cudaFree(b.frontier_count);
    cudaFree(b.rest); cudaFree(b.flags);
    b = b200aabb_tree::BuildScratch();
}
void free_input(b200aabb_tree* t) { cudaFree(t->d_soup); cudaFree(t->d_ids); t->d_soup = nullptr; t->d_ids = nullptr; }

void empty_bbox(double* b) { for (int i = 0; i < 3; ++i) { b[i] = INFINITY; b[3 + i] = -INFINITY; } }

int ensure_stage(b200aabb_tree* t, int slot, size_t in_bytes, size_t out_bytes) {   // (grow only; cudaFree waits for the work in flight)
    if (in_bytes > t->stage_in_bytes[slot]) {
        cudaFree(t->d_stage_in[slot]); t->d_stage_in[slot] = nullptr; t->stage_in_bytes[slot] = 0;
        B2_CUDA(t, cudaMalloc(&t->d_stage_in[slot], in_bytes));
        t->stage_in_bytes[slot] = in_bytes;
    }
    if (out_bytes > t->stage_out_bytes[slot]) {
        cudaFree(t->d_stage_out[slot]); t->d_stage_out[slot] = nullptr; t->stage_out_bytes[slot] = 0;
        B2_CUDA(t, cudaMalloc(&t->d_stage_out[slot], out_bytes));
        t->stage_out_bytes[slot] = out_bytes;
    }
    return B200AABB_OK;
}

void free_warp(b200aabb_tree* t) {
    for (auto& w : t->warp) { cudaFree(w.cursor); cudaFree(w.spill); cudaFree(w.slots64); cudaFree(w.fc_cand); cudaFree(w.fc_ncand); cudaFree(w.fc_overflow); cudaFree(w.fc_count);
                              w = b200aabb_tree::WarpScratch(); }
}
int ensure_warp(b200aabb_tree* t, int slot) {
    if (t->warp_grid == 0) {
        int per_sm = 0, sms = 0;
        B2_CUDA(t, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, closest_point_warp_kernel<false>, 32, 0));
        B2_CUDA(t, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device));
        t->warp_grid = std::max(1, per_sm) * std::max(1, sms);
    }
    if (t->pair_grid == 0) {
        int per_sm = 0, sms = 0;
        B2_CUDA(t, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, closest_point_pair_kernel, 32, 0));
        B2_CUDA(t, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device));
        t->pair_grid = std::max(1, per_sm) * std::max(1, sms);
    }
    static_assert(kPpSpillPerWarp <= kWpSpillCap, "the pair kernel's spill area fits the pool kernel's");
    const size_t grid = (size_t)std::max(t->warp_grid, t->pair_grid);
    auto& w = t->warp[slot];
    if (!w.cursor) B2_CUDA(t, cudaMalloc(&w.cursor, sizeof(unsigned long long)));
    if (!w.spill) B2_CUDA(t, cudaMalloc(&w.spill, grid * kWpSpillCap * sizeof(uint2)));
    if (!w.slots64) B2_CUDA(t, cudaMalloc(&w.slots64, grid * sizeof(WarpSlots64)));
    return B200AABB_OK;
}

int ensure_filter(b200aabb_tree* t, int slot, uint64_t n) {      // candidate lists of the two-pass closest_point kernels (grow only)
    auto& w = t->warp[slot];
    if (n <= w.fc_cap) return B200AABB_OK;
    cudaFree(w.fc_cand); cudaFree(w.fc_ncand); cudaFree(w.fc_overflow); cudaFree(w.fc_count);
    w.fc_cand = nullptr; w.fc_ncand = nullptr; w.fc_overflow = nullptr; w.fc_count = nullptr; w.fc_cap = 0;
    B2_CUDA(t, cudaMalloc(&w.fc_cand, n * kFcCand * 4));
    B2_CUDA(t, cudaMalloc(&w.fc_ncand, n));
    B2_CUDA(t, cudaMalloc(&w.fc_overflow, n * 4));
    B2_CUDA(t, cudaMalloc(&w.fc_count, sizeof(unsigned long long)));
    w.fc_cap = n;
    return B200AABB_OK;
}

void free_sort(b200aabb_tree* t) {
    for (auto& sc : t->sort) {
        cudaFree(sc.keys[0]); cudaFree(sc.keys[1]); cudaFree(sc.vals[0]); cudaFree(sc.vals[1]); cudaFree(sc.ws);
        sc = b200aabb_tree::SortScratch();
    }
}
int ensure_sort(b200aabb_tree* t, int slot, uint64_t n) {
    auto& sc = t->sort[slot];
    if (n <= sc.cap) return B200AABB_OK;
    cudaFree(sc.keys[0]); cudaFree(sc.keys[1]); cudaFree(sc.vals[0]); cudaFree(sc.vals[1]); cudaFree(sc.ws);
    sc = b200aabb_tree::SortScratch();
    for (int k = 0; k < 2; ++k) { B2_CUDA(t, cudaMalloc(&sc.keys[k], n * 8)); B2_CUDA(t, cudaMalloc(&sc.vals[k], n * 4)); }
    B2_CUDA(t, cudaMalloc(&sc.ws, rs_workspace_bytes(n)));
    sc.cap = n;
    return B200AABB_OK;
}

enum QueryShape { kPointQuery, kLineQuery, kLineVecQuery /* Line_3: point + direction vector */, kPlaneQuery /* not sorted */, kTriangleQuery /* sorted by the first vertex */ };

// coherence order of a batch resident on the device: keys -> onesweep sort -> permutation (nullptr when sorting is off)
int coherence_order(b200aabb_tree* t, int slot, cudaStream_t st, const void* d_in, uint64_t n, QueryShape shape, const uint32_t** perm) {
    *perm = nullptr;
    if (!t->opt_sort_queries || n < 4096 || n >= (1ull << 30) || shape == kPlaneQuery) return B200AABB_OK;   // (the sort's look-back words carry 30-bit prefixes)
    int rc = ensure_sort(t, slot, n);
    if (rc) return rc;
    auto& sc = t->sort[slot];
    const unsigned blocks = (unsigned)((n + 255) / 256);
    int passes;
    if (shape == kPointQuery || shape == kTriangleQuery) { point_keys_kernel<<<blocks, 256, 0, st>>>((const double*)d_in, shape == kTriangleQuery ? 9u : 3u, n, t->d_rootbox, sc.keys[0], sc.vals[0]); passes = 4; }
    else { line_keys_kernel<<<blocks, 256, 0, st>>>((const double*)d_in, n, t->d_rootbox, sc.keys[0], sc.vals[0], shape == kLineVecQuery ? 1 : 0); passes = 6; }
    t->stats.kernel_launches += 1;
    const int cur = radix_sort_pairs(sc.keys, sc.vals, n, passes, sc.ws, st, &t->stats.kernel_launches);
    *perm = sc.vals[cur];
    return B200AABB_OK;
}

TreeRef tree_ref(const b200aabb_tree* t, const uint32_t* perm) { return TreeRef{t->d_nodes, t->d_tris, t->root, t->mag, perm, t->d_counters}; }

// scratch of the build and of the refit for n primitives, kept with the handle (grow only, given back when four times too large)
int ensure_build_scratch(b200aabb_tree* t, uint64_t n) {
    auto& bs = t->bs;
    if (n <= bs.cap && n >= bs.cap / 4) return B200AABB_OK;
    B2_CUDA(t, cudaStreamSynchronize(t->stream));
    free_build_scratch(t);
    const uint64_t frontier_cap = (n + kRefitBlock - 1) / kRefitBlock * kRefitFrontierPerBlock;
    cudaError_t e = cudaSuccess;
    for (int s_ = 0; s_ < 2 && e == cudaSuccess; ++s_) e = cudaMalloc(&bs.keys[s_], n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&bs.ws, rs_workspace_bytes(n));
    if (e == cudaSuccess) e = cudaMalloc(&bs.flags, 16);                             // [0] runs of equal PACKED key bits, [1] frontier overflow of the fused path
    if (e == cudaSuccess) e = cudaMalloc(&bs.frontier, frontier_cap * sizeof(RefitEntry));
    if (e == cudaSuccess) e = cudaMalloc(&bs.frontier_count, 2 * sizeof(unsigned));  // [1]: length of `rest`
    if (e == cudaSuccess) e = cudaMalloc(&bs.rest, frontier_cap * 4);               // nodes the fused path leaves to the searches (fewer than its frontier entries)
    if (e != cudaSuccess) { free_build_scratch(t); t->err = std::string("cudaMalloc(build scratch): ") + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? B200AABB_ERR_OUT_OF_MEMORY : B200AABB_ERR_CUDA; }
    bs.cap = n;
    return B200AABB_OK;
}

int do_build(b200aabb_tree* t) {
    B2_CUDA(t, cudaSetDevice(t->device));
    if (t->imported) {   // a replica received from another rank: it IS built, and it has no primitive input to rebuild from
        if (t->built) return B200AABB_OK;
        return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "an imported tree can not be rebuilt: insert primitives first");
    }
    { const int rc_ = sync_pipeline(t); if (rc_) return rc_; }   // (asynchronous HOST batches may still be in flight)
    const uint64_t n = t->n;
    t->built = false; t->root = kNone;
    empty_bbox(t->h_bbox);
    if (n == 0) { free_tree_buffers(t); t->built = true; t->stats.last_build_ms = 0; return B200AABB_OK; }
    if (n >= (1ull << 30)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "at most 2^30 - 1 primitives per tree");
    cudaStream_t st = t->stream;
    auto& bs = t->bs;
    auto cleanup = [&]() {};                                          // (the scratch stays with the handle)
#define B2_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { t->err = std::string(#call) + ": " + cudaGetErrorString(e_); free_build_scratch(t); free_tree_buffers(t); \
        return e_ == cudaErrorMemoryAllocation ? B200AABB_ERR_OUT_OF_MEMORY : B200AABB_ERR_CUDA; } } while (0)
    const uint64_t refit_blocks = (n + kRefitBlock - 1) / kRefitBlock, frontier_cap = refit_blocks * kRefitFrontierPerBlock, frontier_use = refit_blocks * (uint64_t)t->opt_debug_refit_frontier;
    { const int rc_ = ensure_build_scratch(t, n); if (rc_) { free_tree_buffers(t); return rc_; } }
    if (n > t->tree_cap || n < t->tree_cap / 4 || !t->d_tris) {
        B2_TRY(cudaStreamSynchronize(st));
        free_tree_buffers(t);
        B2_TRY(cudaMalloc(&t->d_leaf_parent, n * 4));
        B2_TRY(cudaMalloc(&t->d_nodes, std::max<uint64_t>(n - 1, 1) * sizeof(BvhNode)));
        B2_TRY(cudaMalloc(&t->d_tris, n * sizeof(TriRec)));
        B2_TRY(cudaMalloc(&t->d_trisf, n * sizeof(TriF)));
        t->tree_cap = n;
    }
    uint64_t** const keys = bs.keys;
    uint32_t** const vals = bs.vals;
    void* const ws = bs.ws; RefitEntry* const frontier = bs.frontier; unsigned* const frontier_count = bs.frontier_count; unsigned long long* const d_flags = bs.flags;
    uint32_t* const rest = bs.rest;
    uint32_t* const leaf_parent = t->d_leaf_parent;
    // B200AABB_BUILD_TIMING=1: per-stage device times on stderr (events between the stages; diagnostic only)
    const bool stage_timing = getenv("B200AABB_BUILD_TIMING") != nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* name) { if (stage_timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); marks.push_back({name, e}); } };
    const unsigned blocks = (unsigned)((n + 255) / 256), pblocks = std::min<unsigned>(blocks, 148 * 8);
    const int ib = index_bits_for(n);                                 // PACKED words: index in the low ib bits, 64 - ib key bits above
    const int packed_passes = std::min(8, (64 - ib + 7) / 8);
    unsigned long long dups = 0;
    EmitArgs emit{};
    B2_TRY(cudaEventRecord(t->ev_a, st));
    mark("start");
    if (!t->bounds_valid) {
        B2_TRY(cudaMemsetAsync(t->d_bounds, 0xFF, 3 * 8, st));
        B2_TRY(cudaMemsetAsync(t->d_bounds + 3, 0x00, 3 * 8, st));
        scene_bounds_kernel<<<pblocks, 256, 0, st>>>(t->d_soup, n, t->d_bounds);
        t->stats.kernel_launches += 1;
        t->bounds_valid = true;
    }
    mark("scene_bounds");
    // Two attempts at most: PACKED sort words first (build.cuh::pack_sort_word); if more than 1/16 of the primitives start a run of four with equal kept key
    // bits (a scene whose primitives crowd a tiny part of its box), once more with full 63-bit keys + values.
    // After the sort, the FUSED path: build_local_kernel (leaf records, topology and boxes of all nodes inside a block of leaves, bottom-up),
    // hierarchy_kernel for the few nodes that span blocks, refit_global_kernel for their boxes.  Its frontier list filling up (a degenerate,
    // chain-like tree) sends the build down the CLASSIC path: hierarchy_kernel for every node, refit_local_kernel, refit_global_kernel.
    bool classic = t->opt_classic_build != 0;
    for (int attempt = t->opt_exact_keys ? 1 : 0; attempt < 2; ++attempt) {
        const bool packed = attempt == 0;
        int cur;
        if (packed) {
            rs_begin(ws, n, st);
            morton_kernel<true><<<pblocks, 256, 0, st>>>(t->d_soup, n, t->d_bounds, keys[0], nullptr, ib, packed_passes, rs_hist_of(ws));
            mark("morton");
            t->stats.kernel_launches += 1;
            cur = rs_sort_passes(keys, nullptr, n, ib, packed_passes, ws, st, &t->stats.kernel_launches);
        } else {
            if (bs.vals_cap < n) { for (int s_ = 0; s_ < 2; ++s_) { cudaFree(vals[s_]); vals[s_] = nullptr; B2_TRY(cudaMalloc(&vals[s_], n * 4)); } bs.vals_cap = n; }
            morton_kernel<false><<<pblocks, 256, 0, st>>>(t->d_soup, n, t->d_bounds, keys[0], vals[0], 0, 0, nullptr);
            mark("morton");
            t->stats.kernel_launches += 1;
            cur = radix_sort_pairs(keys, vals, n, 8, ws, st, &t->stats.kernel_launches);
        }
        mark("sort");
        emit = EmitArgs{t->d_soup, t->d_ids, packed ? nullptr : vals[cur], keys[cur], packed ? ib : 0, t->d_trisf};
        unsigned long long flags[2] = {0, 0};
        while (true) {
            B2_TRY(cudaMemsetAsync(d_flags, 0, 16, st));
            B2_TRY(cudaMemsetAsync(frontier_count, 0, 2 * sizeof(unsigned), st));
            if (!classic) {
                build_local_kernel<<<(unsigned)refit_blocks, kRefitBlock, 0, st>>>(keys[cur], t->d_tris, n, t->d_nodes, leaf_parent, frontier, frontier_count, (unsigned)frontier_use,
                                                                                   emit, rest, frontier_count + 1, d_flags);
                mark("leaves+tree");
                if (n > 1) hierarchy_kernel<<<(unsigned)((frontier_use + 255) / 256) + 1, 256, 0, st>>>(keys[cur], n, t->d_nodes, leaf_parent, 0, nullptr, rest, frontier_count + 1, (unsigned)frontier_use);
                mark("hierarchy_rest");
                t->stats.kernel_launches += n > 1 ? 2 : 1;
            } else {
                if (n > 1) hierarchy_kernel<<<blocks, 256, 0, st>>>(keys[cur], n, t->d_nodes, leaf_parent, packed ? ib : 0, d_flags, nullptr, nullptr, 0);
                mark("hierarchy");
                refit_local_kernel<true><<<(unsigned)refit_blocks, kRefitBlock, 0, st>>>(t->d_tris, n, t->d_nodes, leaf_parent, frontier, frontier_count, (unsigned)frontier_use, emit);
                mark("leaves+refit");
                t->stats.kernel_launches += n > 1 ? 2 : 1;
            }
            if (n > 1) { refit_global_kernel<<<(unsigned)((frontier_cap + 255) / 256), 256, 0, st>>>(t->d_nodes, leaf_parent, frontier, frontier_count, (unsigned)frontier_use); t->stats.kernel_launches += 1; }
            mark("refit_global");
            B2_TRY(cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, st));
            B2_TRY(cudaStreamSynchronize(st));
            if (!classic && flags[1]) { classic = true; continue; }      // frontier overflow: the classic path finishes such climbs in place
            break;
        }
        dups = flags[0];
        if (packed && dups > n / 16) continue;
        t->last_build_packed = packed ? 1 : 0;
        t->last_build_dups = packed ? dups : 0;
        t->last_build_classic = classic ? 1 : 0;
        break;
    }
    root_box_kernel<<<1, 32, 0, st>>>(t->d_bounds, t->d_rootbox);
    mark("root_box");
    t->stats.kernel_launches += 2;
    B2_TRY(cudaEventRecord(t->ev_b, st));
    B2_TRY(cudaMemcpyAsync(t->h_bbox, t->d_rootbox, 48, cudaMemcpyDeviceToHost, st));
    B2_TRY(cudaStreamSynchronize(st));
    B2_TRY(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, t->ev_a, t->ev_b);
    t->stats.last_build_ms = ms;
    for (size_t i = 1; i < marks.size(); ++i) { float d = 0; cudaEventElapsedTime(&d, marks[i - 1].second, marks[i].second); fprintf(stderr, "[b200aabb build] %-14s %8.3f ms\n", marks[i].first, d); }
    if (stage_timing) fprintf(stderr, "[b200aabb build] %s sort words, %d passes, %llu runs of >= 4 equal key bits, %s path\n", t->last_build_packed ? "packed" : "full (key, index)", t->last_build_packed ? packed_passes : 8, dups, t->last_build_classic ? "classic" : "fused");
    for (auto& m : marks) cudaEventDestroy(m.second);
    t->root = n == 1 ? (kLeafBit | 0u) : 0u;
    t->mag = 0.0;
    for (int a = 0; a < 6; ++a) t->mag = std::max(t->mag, fabs(t->h_bbox[a]));
    t->built = true;
    cleanup();
#undef B2_TRY
    return B200AABB_OK;
}

int ensure_built(b200aabb_tree* t) { return t->built ? B200AABB_OK : do_build(t); }


// Runs `launch(stream, d_in, d_outs, count, perm)` over the batch (perm = coherence order of the chunk, or nullptr).  HOST pointers: chunked, double-buffered
// H2D -> kernel -> D2H pipeline on two streams; DEVICE pointers: one launch on the handle's stream, no synchronisation.
struct OutSpec { void* host; size_t elem_bytes; size_t offset; };   // one result array: host pointer, bytes per query, offset in staging

int sync_pipeline(b200aabb_tree* t) {
    B2_CUDA(t, cudaStreamSynchronize(t->s_up));
    B2_CUDA(t, cudaStreamSynchronize(t->s_run));
    B2_CUDA(t, cudaStreamSynchronize(t->s_down));
    return B200AABB_OK;
}

// zeroes the statistics counters of a batch ([2], the incomplete-traversal flags, is sticky while asynchronous batches are in flight)
int reset_counters(b200aabb_tree* t, cudaStream_t st) {
    if (!t->async_pending) { B2_CUDA(t, cudaMemsetAsync(t->d_counters, 0, 8 * sizeof(unsigned long long), st)); return B200AABB_OK; }
    B2_CUDA(t, cudaMemsetAsync(t->d_counters, 0, 2 * sizeof(unsigned long long), st));
    B2_CUDA(t, cudaMemsetAsync(t->d_counters + 3, 0, 5 * sizeof(unsigned long long), st));
    return B200AABB_OK;
}

// waits for every batch in flight on the handle and reports a traversal that could not finish
int drain(b200aabb_tree* t) {
    { const int rc_ = sync_pipeline(t); if (rc_) return rc_; }
    B2_CUDA(t, cudaStreamSynchronize(t->stream));
    B2_CUDA(t, cudaGetLastError());
    t->async_pending = false;
    if (!t->trace.empty()) {
        for (auto& r : t->trace) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, t->trace[0].e[0], r.e[0]); cudaEventElapsedTime(&b, t->trace[0].e[0], r.e[1]);
            cudaEventElapsedTime(&c, t->trace[0].e[0], r.e[2]); cudaEventElapsedTime(&d, t->trace[0].e[0], r.e[3]);
            fprintf(stderr, "[b200aabb pipe] slot %d %s %8llu  start %8.3f  h2d-done %8.3f  kernels-done %8.3f  d2h-done %8.3f\n", r.slot, r.shape == kPointQuery ? "points" : "lines ",
                    (unsigned long long)r.cnt, a, b, c, d);
        }
        for (auto& r : t->trace) for (auto& e : r.e) cudaEventDestroy(e);
        t->trace.clear();
    }
    unsigned long long incomplete = 0;
    B2_CUDA(t, cudaMemcpy(&incomplete, t->d_counters + 2, sizeof(incomplete), cudaMemcpyDeviceToHost));
    if (incomplete) {
        cudaMemset(t->d_counters + 2, 0, sizeof(incomplete));
        return fail(t, (incomplete & 4) ? B200AABB_ERR_INVALID_ARGUMENT : B200AABB_ERR_INCOMPLETE,
                    (incomplete & 4) ? "offsets disagree with the number of intersected primitives" : "traversal stack overflow: results incomplete");
    }
    return B200AABB_OK;
}

template <class Launch>
int run_batch(b200aabb_tree* t, const void* in, size_t in_elem_bytes, uint64_t n, b200aabb_memspace space, QueryShape shape, std::vector<OutSpec> outs, Launch launch) {
    B2_CUDA(t, cudaSetDevice(t->device));
    if (space == B200AABB_DEVICE) {
        int rc = reset_counters(t, t->stream);
        if (rc) return rc;
        B2_CUDA(t, cudaEventRecord(t->ev_a, t->stream));
        std::vector<void*> ptrs; for (auto& o : outs) ptrs.push_back(o.host);
        const uint32_t* perm = nullptr;
        rc = coherence_order(t, kDeviceScratch, t->stream, in, n, shape, &perm);
        if (rc) return rc;
        t->cur_slot = kDeviceScratch; t->cur_offset = 0;
        launch(t->stream, in, ptrs.data(), n, perm);
        t->stats.kernel_launches += 1;
        B2_CUDA(t, cudaEventRecord(t->ev_b, t->stream));
        B2_CUDA(t, cudaGetLastError());
        t->query_timing_pending = true;
        return B200AABB_OK;
    }
    // HOST pointers.  Ray / segment batches are cut into chunks that flow through upload -> sort -> traverse -> download on kLineSlots
    // staging slots, so that the copies of one chunk hide behind the kernels of the others (such batches are copy bound: ~1 ms of PCIe per Mi rays
    // each way, 0.6 ms of kernel).  A point batch stays WHOLE: its queries are sorted into one coherence order, and a chunk of an unordered
    // batch covers the whole scene at a fraction of the density — measured (scripts/tune_e2e.py, profiles/r2_05): four chunks of 2 Mi
    // points take 4 x 8 ms against 23.8 ms for the batch in one launch.  Its copies hide behind OTHER batches instead (B200AABB_OPT_ASYNC_HOST).
    const bool points = shape == kPointQuery;
    const uint64_t chunk = chunk_for(n, t->opt_host_parts > 0 ? t->opt_host_parts : (points ? 1 : 8), points ? kPointChunkMax : kChunkMax);
    size_t out_bytes = 0;
    for (auto& o : outs) { o.offset = out_bytes; out_bytes += ((o.host ? o.elem_bytes * chunk : 0) + 255) / 256 * 256; }
    int rc = B200AABB_OK;
    B2_CUDA(t, cudaEventRecord(t->ev_a, t->stream));
    B2_CUDA(t, cudaStreamWaitEvent(t->s_run, t->ev_a, 0));               // (whatever the caller enqueued on the handle's stream, e.g. the build)
    rc = reset_counters(t, t->s_run);
    if (rc) return rc;
    static const bool tracing = getenv("B200AABB_PIPE_TRACE") != nullptr;
    for (uint64_t b = 0; b < n; b += chunk) {
        const uint64_t cnt = std::min<uint64_t>(chunk, n - b);
        int slot;
        if (points) { slot = kLineSlots + t->next_point_slot; t->next_point_slot = (t->next_point_slot + 1) % kPointSlots; }
        else { slot = t->next_line_slot; t->next_line_slot = (t->next_line_slot + 1) % kLineSlots; }
        rc = ensure_stage(t, slot, in_elem_bytes * chunk, std::max<size_t>(out_bytes, 256));
        if (rc) return rc;
        b200aabb_tree::TraceRec tr;
        if (tracing) { for (auto& e : tr.e) cudaEventCreate(&e); tr.slot = slot; tr.cnt = cnt; tr.shape = (int)shape; }
        // upload: once the kernels that last read this slot's input have finished
        B2_CUDA(t, cudaStreamWaitEvent(t->s_up, t->ev_run[slot], 0));
        if (tracing) cudaEventRecord(tr.e[0], t->s_up);
        B2_CUDA(t, cudaMemcpyAsync(t->d_stage_in[slot], (const char*)in + b * in_elem_bytes, cnt * in_elem_bytes, cudaMemcpyHostToDevice, t->s_up));
        B2_CUDA(t, cudaEventRecord(t->ev_up[slot], t->s_up));
        if (tracing) cudaEventRecord(tr.e[1], t->s_up);
        // kernels: once the chunk is there and the slot's previous results have left
        B2_CUDA(t, cudaStreamWaitEvent(t->s_run, t->ev_up[slot], 0));
        B2_CUDA(t, cudaStreamWaitEvent(t->s_run, t->ev_down[slot], 0));
        std::vector<void*> ptrs;
        for (auto& o : outs) ptrs.push_back(o.host ? (char*)t->d_stage_out[slot] + o.offset : nullptr);
        const uint32_t* perm = nullptr;
        rc = coherence_order(t, kHostScratch, t->s_run, t->d_stage_in[slot], cnt, shape, &perm);
        if (rc) return rc;
        t->cur_slot = slot; t->cur_offset = b;
        launch(t->s_run, t->d_stage_in[slot], ptrs.data(), cnt, perm);
        t->stats.kernel_launches += 1;
        B2_CUDA(t, cudaEventRecord(t->ev_run[slot], t->s_run));
        if (tracing) cudaEventRecord(tr.e[2], t->s_run);
        // download
        B2_CUDA(t, cudaStreamWaitEvent(t->s_down, t->ev_run[slot], 0));
        for (auto& o : outs)
            if (o.host) B2_CUDA(t, cudaMemcpyAsync((char*)o.host + b * o.elem_bytes, (char*)t->d_stage_out[slot] + o.offset, cnt * o.elem_bytes, cudaMemcpyDeviceToHost, t->s_down));
        B2_CUDA(t, cudaEventRecord(t->ev_down[slot], t->s_down));
        if (tracing) { cudaEventRecord(tr.e[3], t->s_down); t->trace.push_back(tr); }
    }
    if (t->opt_async_host) {   // the caller keeps its buffers alive and calls b200aabb_synchronize before it reads the results
        t->async_pending = true;
        t->query_timing_pending = false;
        return B200AABB_OK;
    }
    B2_CUDA(t, cudaEventRecord(t->ev_b, t->s_down));
    rc = drain(t);   // a traversal that could not finish is an error of THIS call, never a silent partial result
    t->query_timing_pending = true;
    return rc;
}

unsigned grid_for(uint64_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------------
// launches the intersection kernel of a query kind (Segment_3 / Ray_3 / Line_3: 6 doubles per query; Plane_3: 4) in one of the modes of bvh.cuh
template <int MODE>
static void launch_intersect(b200aabb_tree* t, int kind, cudaStream_t st, const double* in, uint64_t cnt, const uint32_t* perm, uint8_t* out8, uint32_t* out32,
                             const uint64_t* offsets, uint32_t* list) {
    if constexpr (MODE != kList) {
        if (t->opt_ray_kernel == 1 && t->n >= 2 && kind != B200AABB_PLANE && kind != B200AABB_TRIANGLE && t->d_ray_cursor) {
            unsigned long long* cursor = t->d_ray_cursor + (st == t->s_run ? kHostScratch : kDeviceScratch);   // (HOST batches run on s_run, DEVICE batches on the caller's stream)
            cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st);
            const unsigned sgrid = (unsigned)std::min<uint64_t>((uint64_t)t->seg_grid, (cnt + 127) / 128);
            switch (kind) {
                case B200AABB_RAY: intersect_stream_kernel<kRayQ, MODE><<<sgrid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, cursor, out8, out32); break;
                case B200AABB_LINE: intersect_stream_kernel<kLineQ, MODE><<<sgrid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, cursor, out8, out32); break;
                default: intersect_stream_kernel<kSegmentQ, MODE><<<sgrid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, cursor, out8, out32); break;
            }
            return;
        }
    }
    const unsigned grid = grid_for(cnt, 128);
    switch (kind) {
        case B200AABB_RAY: intersect_kernel<kRayQ, MODE><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, out8, out32, offsets, list); break;
        case B200AABB_LINE: intersect_kernel<kLineQ, MODE><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, out8, out32, offsets, list); break;
        case B200AABB_PLANE: plane_kernel<MODE><<<grid, 128, 0, st>>>(tree_ref(t, nullptr), in, cnt, out8, out32, offsets, list); break;
        case B200AABB_TRIANGLE:
            if (t->segments) triangle_kernel<MODE, true><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, 9u, cnt, out8, out32, offsets, list);
            else triangle_kernel<MODE, false><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, 9u, cnt, out8, out32, offsets, list);
            break;
        default: intersect_kernel<kSegmentQ, MODE><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, out8, out32, offsets, list); break;
    }
}
// a box that contains the image of the box b (xmin ymin zmin xmax ymax zmax) under the affine map xf (12 doubles; nullptr = identity), widened by the rounding
// of the images of the primitives inside (as bvh.cuh::AffXf does per node); false = the map is not finite
static bool image_of_box(const double* xf, const double* b, double* out) {
    if (!xf) { for (int k = 0; k < 6; ++k) out[k] = b[k]; return true; }
    for (int k = 0; k < 3; ++k) {
        const double* r = xf + 4 * k;
        const double c[3] = {0.5 * (b[0] + b[3]), 0.5 * (b[1] + b[4]), 0.5 * (b[2] + b[5])}, h[3] = {0.5 * (b[3] - b[0]), 0.5 * (b[4] - b[1]), 0.5 * (b[5] - b[2])};
        const double ctr = ((r[0] * c[0] + r[1] * c[1]) + r[2] * c[2]) + r[3];
        const double ext = (fabs(r[0]) * h[0] + fabs(r[1]) * h[1]) + fabs(r[2]) * h[2];
        const double mag = ((fabs(r[0]) * (fabs(c[0]) + h[0]) + fabs(r[1]) * (fabs(c[1]) + h[1])) + fabs(r[2]) * (fabs(c[2]) + h[2])) + fabs(r[3]);
        if (!(ctr == ctr) || !(ext == ext)) return false;
        out[k] = ctr - (ext + 1e-14 * mag); out[3 + k] = ctr + (ext + 1e-14 * mag);
    }
    return true;
}
// a tree of Segment_3 primitives answers closest_point / squared_distance and the id queries with Plane_3 and Triangle_3 (what the reference's kernel can
// intersect with a Segment_3 datum robustly in 3D); every other entry point needs triangles
static int need_triangles(b200aabb_tree* t, const char* what) {
    if (!t->segments) return B200AABB_OK;
    t->err = std::string(what) + ": not offered on a tree of Segment_3 primitives";
    return B200AABB_ERR_INVALID_ARGUMENT;
}
static bool kind_fits_primitives(const b200aabb_tree* t, int kind) { return !t->segments || kind == B200AABB_PLANE || kind == B200AABB_TRIANGLE; }
static bool valid_kind(int kind) { return kind >= B200AABB_SEGMENT && kind <= B200AABB_TRIANGLE; }          // the id queries
static bool valid_object_kind(int kind) { return kind >= B200AABB_SEGMENT && kind <= B200AABB_PLANE; }    // the intersection OBJECTS (no Triangle_3 queries)
static size_t query_bytes(int kind) { return kind == B200AABB_PLANE ? 32 : (kind == B200AABB_TRIANGLE ? 72 : 48); }
static QueryShape query_shape(int kind) {
    return kind == B200AABB_PLANE ? kPlaneQuery : (kind == B200AABB_LINE ? kLineVecQuery : (kind == B200AABB_TRIANGLE ? kTriangleQuery : kLineQuery));
}
template <int MODE>
static void launch_objects(b200aabb_tree* t, int kind, cudaStream_t st, const double* in, uint64_t cnt, const uint32_t* perm, uint32_t* out32, const uint64_t* offsets,
                           uint32_t* ids, uint8_t* kinds, double* pts, uint32_t* prim, uint8_t* flags, double* pt, double* pt2) {
    const unsigned grid = grid_for(cnt, 128);
    switch (kind) {
        case B200AABB_SEGMENT: ray_objects_kernel<MODE, kSegmentQ><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, out32, offsets, ids, kinds, pts, prim, flags, pt, pt2); break;
        case B200AABB_LINE: ray_objects_kernel<MODE, kLineQ><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, out32, offsets, ids, kinds, pts, prim, flags, pt, pt2); break;
        case B200AABB_PLANE: ray_objects_kernel<MODE, kPlaneObjects><<<grid, 128, 0, st>>>(tree_ref(t, nullptr), in, cnt, out32, offsets, ids, kinds, pts, prim, flags, pt, pt2); break;
        default: ray_objects_kernel<MODE, kRayQ><<<grid, 128, 0, st>>>(tree_ref(t, perm), in, cnt, out32, offsets, ids, kinds, pts, prim, flags, pt, pt2); break;
    }
}


extern "C" {

const char* b200aabb_version(void) { return "b200aabb 0.1 sm_100a"; }

int b200aabb_debug_predicate(int device, int which, const double* records, int doubles_per_record, uint64_t n, int* result, int* used_exact, double* values) {
    static const int need[] = {12, 15, 9, 9, 15, 15, 15, 12, 12};
    if (which < 0 || which > 8 || !records || !result || doubles_per_record < need[which]) return B200AABB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B200AABB_ERR_NO_DEVICE;
    if (n == 0) return B200AABB_OK;
    double* d_in = nullptr; int* d_out = nullptr; int* d_ex = nullptr; double* d_val = nullptr;
    bool ok = cudaMalloc(&d_in, n * doubles_per_record * 8) == cudaSuccess && cudaMalloc(&d_out, n * 4) == cudaSuccess && cudaMalloc(&d_ex, n * 4) == cudaSuccess &&
              cudaMalloc(&d_val, n * 64) == cudaSuccess;
    ok = ok && cudaMemcpy(d_in, records, n * doubles_per_record * 8, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemset(d_val, 0, n * 64) == cudaSuccess;
    if (ok) {
        predicate_kernel<<<(unsigned)((n + 127) / 128), 128>>>(which, d_in, doubles_per_record, n, d_out, d_ex, d_val);
        ok = cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(result, d_out, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok && used_exact) ok = cudaMemcpy(used_exact, d_ex, n * 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok && values) ok = cudaMemcpy(values, d_val, n * 64, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_ex); cudaFree(d_val);
    return ok ? B200AABB_OK : B200AABB_ERR_CUDA;
}

int b200aabb_create(int device, b200aabb_tree** out) {
    if (!out) return B200AABB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return B200AABB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return B200AABB_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return B200AABB_ERR_CUDA;
    b200aabb_tree* t = new b200aabb_tree();
    t->device = device;
    memset(&t->stats, 0, sizeof(t->stats));
    empty_bbox(t->h_bbox);
    bool ok = cudaMalloc(&t->d_rootbox, 48) == cudaSuccess && cudaMalloc(&t->d_bounds, 48) == cudaSuccess && cudaMalloc(&t->d_counters, 8 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaEventCreate(&t->ev_a) == cudaSuccess && cudaEventCreate(&t->ev_b) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&t->s_up, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&t->s_run, cudaStreamNonBlocking) == cudaSuccess &&
         cudaStreamCreateWithFlags(&t->s_down, cudaStreamNonBlocking) == cudaSuccess;
    for (int s = 0; s < kPipeSlots && ok; ++s)
        ok = cudaEventCreateWithFlags(&t->ev_up[s], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&t->ev_run[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&t->ev_down[s], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&t->d_ray_cursor, kScratchSlots * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemset(t->d_counters, 0, 8 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&t->d_witness, 3 * sizeof(unsigned)) == cudaSuccess;   // (drain() may read them before the first query has reset them)
    if (ok) {
        int per_sm = 0;
        ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, first_intersection_stream_kernel, 128, 0) == cudaSuccess;
        t->ray_grid = std::max(1, per_sm) * prop.multiProcessorCount;
        ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, intersect_stream_kernel<kSegmentQ, kAnyHit>, 128, 0) == cudaSuccess;
        t->seg_grid = std::max(1, per_sm) * prop.multiProcessorCount;
    }
    if (!ok) { b200aabb_destroy(t); return B200AABB_ERR_CUDA; }
    t->built = true;   // an empty tree needs no build
    *out = t;
    return B200AABB_OK;
}

int b200aabb_destroy(b200aabb_tree* t) {
    if (!t) return B200AABB_OK;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();
    free_tree_buffers(t); free_build_scratch(t); free_input(t); free_sort(t); free_warp(t);
    cudaFree(t->d_rootbox); cudaFree(t->d_bounds); cudaFree(t->d_counters); cudaFree(t->d_ray_cursor); cudaFree(t->d_witness);
    for (int s = 0; s < kPipeSlots; ++s) {
        cudaFree(t->d_stage_in[s]); cudaFree(t->d_stage_out[s]); cudaFree(t->d_stage_aux[s]);
        if (t->ev_up[s]) cudaEventDestroy(t->ev_up[s]);
        if (t->ev_run[s]) cudaEventDestroy(t->ev_run[s]);
        if (t->ev_down[s]) cudaEventDestroy(t->ev_down[s]);
    }
    for (cudaStream_t st : {t->s_up, t->s_run, t->s_down}) if (st) cudaStreamDestroy(st);
    if (t->ev_a) cudaEventDestroy(t->ev_a);
    if (t->ev_b) cudaEventDestroy(t->ev_b);
    delete t;
    return B200AABB_OK;
}

int b200aabb_set_option(b200aabb_tree* t, int option, int64_t value) { B2_LOCK(t);
    if (!t) return B200AABB_ERR_INVALID_ARGUMENT;
    switch (option) {
        case B200AABB_OPT_SORT_QUERIES: t->opt_sort_queries = value != 0; return B200AABB_OK;
        case B200AABB_OPT_CLOSEST_KERNEL: t->opt_closest_kernel = (int)std::min<int64_t>(3, std::max<int64_t>(0, value)); return B200AABB_OK;
        case B200AABB_OPT_ASYNC_HOST: t->opt_async_host = value != 0; return B200AABB_OK;
        case B200AABB_OPT_RAY_KERNEL: t->opt_ray_kernel = value != 0; return B200AABB_OK;
        case B200AABB_OPT_HOST_PARTS: t->opt_host_parts = (int)std::min<int64_t>(64, std::max<int64_t>(0, value)); return B200AABB_OK;
        case 103: t->opt_classic_build = value != 0; if (!t->imported) t->built = t->n == 0; return B200AABB_OK;
        case 102: t->opt_exact_keys = value != 0; if (!t->imported) t->built = t->n == 0; return B200AABB_OK;
        case 101: t->opt_debug_refit_frontier = (int)std::min<int64_t>(kRefitFrontierPerBlock, std::max<int64_t>(0, value)); if (!t->imported) t->built = t->n == 0; return B200AABB_OK;
        case B200AABB_OPT_WARP_EXACT_GO: t->opt_warp_exact_go = (int)std::min<int64_t>(kWpExactCap - 32, std::max<int64_t>(0, value)); return B200AABB_OK;
        default: return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "unknown option");
    }
}
int b200aabb_set_stream(b200aabb_tree* t, void* s) { B2_LOCK(t); if (!t) return B200AABB_ERR_INVALID_ARGUMENT; t->stream = (cudaStream_t)s; return B200AABB_OK; }
const char* b200aabb_last_error(const b200aabb_tree* t) { return t ? t->err.c_str() : "null handle"; }

int b200aabb_get_stats(const b200aabb_tree* tc, b200aabb_stats* out) { B2_LOCK(tc);
    b200aabb_tree* t = const_cast<b200aabb_tree*>(tc);
    if (!t || !out) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_CUDA(t, cudaSetDevice(t->device));
    if (t->query_timing_pending) {
        B2_CUDA(t, cudaEventSynchronize(t->ev_b));
        float ms = 0; cudaEventElapsedTime(&ms, t->ev_a, t->ev_b);
        t->stats.last_query_total_ms = ms; t->stats.last_query_ms = ms;
        unsigned long long c[8];
        B2_CUDA(t, cudaMemcpy(c, t->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        t->stats.exact_fallbacks = c[0]; t->stats.ties = c[1]; t->stats.nodes_visited = c[3]; t->stats.leaves_tested = c[4]; t->stats.near_ties = c[5]; t->stats.leaves_culled = c[6];
        if (c[2]) t->err = (c[2] & 4) ? "offsets disagree with the number of intersected primitives" : (c[2] & 2) ? "closest_point work pool inconsistency: results incomplete" : "traversal stack overflow: results incomplete";
        t->query_timing_pending = false;
        if (c[2]) { *out = t->stats; return (c[2] & 4) ? B200AABB_ERR_INVALID_ARGUMENT : B200AABB_ERR_INCOMPLETE; }
    }
    *out = t->stats;
    return B200AABB_OK;
}

int b200aabb_synchronize(b200aabb_tree* t) {
    if (!t) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_LOCK(t);
    B2_CUDA(t, cudaSetDevice(t->device));
    return drain(t);
}

int b200aabb_clear(b200aabb_tree* t) { B2_LOCK(t);
    if (!t) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_CUDA(t, cudaSetDevice(t->device));
    { const int rc_ = sync_pipeline(t); if (rc_) return rc_; }   // (asynchronous HOST batches may still be in flight)
    B2_CUDA(t, cudaStreamSynchronize(t->stream));
    free_tree_buffers(t); free_build_scratch(t); free_input(t);
    t->n = 0; t->built = true; t->imported = false; t->segments = false; t->bounds_valid = false; empty_bbox(t->h_bbox);
    return B200AABB_OK;
}

int b200aabb_set_triangles(b200aabb_tree* t, const double* tris, const uint32_t* ids, uint64_t n, b200aabb_memspace space) { B2_LOCK(t);
    if (!t || (n && !tris)) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = b200aabb_clear(t);
    if (rc) return rc;
    if (n == 0) return B200AABB_OK;
    const cudaMemcpyKind kind = space == B200AABB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    B2_CUDA(t, cudaMalloc(&t->d_soup, n * 72 + kSoupSlack));
    B2_CUDA(t, cudaMemcpyAsync(t->d_soup, tris, n * 72, kind, t->stream));
    if (ids) { B2_CUDA(t, cudaMalloc(&t->d_ids, n * 4)); B2_CUDA(t, cudaMemcpyAsync(t->d_ids, ids, n * 4, kind, t->stream)); }
    B2_CUDA(t, cudaStreamSynchronize(t->stream));
    t->n = n; t->built = false;
    return B200AABB_OK;
}

// Segment_3 primitives (AABB_segment_primitive_3 over a range of segments, AABB_halfedge_graph_segment_primitive over the edges of a graph): each is kept as
// the record (source, target, target), so the build, the boxes and the binary32 pre-tests see the segment as a flat triangle; the leaf routines that must differ
// (projection, Triangle_3 queries) are selected by t->segments.
__global__ void __launch_bounds__(256) expand_segments_kernel(const double* __restrict__ segs, uint64_t n, double* __restrict__ soup) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* s = segs + 6 * i;
    double* o = soup + 9 * i;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; o[4] = s[4]; o[5] = s[5]; o[6] = s[3]; o[7] = s[4]; o[8] = s[5];
}
int b200aabb_set_segments(b200aabb_tree* t, const double* segs, const uint32_t* ids, uint64_t n, b200aabb_memspace space) { B2_LOCK(t);
    if (!t || (n && !segs)) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = b200aabb_clear(t);
    if (rc) return rc;
    t->segments = true;
    if (n == 0) return B200AABB_OK;
    const cudaMemcpyKind kind = space == B200AABB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    double* d_in = nullptr;
    B2_CUDA(t, cudaMalloc(&t->d_soup, n * 72 + kSoupSlack));
    if (space == B200AABB_HOST) {
        if (cudaMalloc(&d_in, n * 48) != cudaSuccess) { free_input(t); t->err = "cudaMalloc(segments)"; return B200AABB_ERR_OUT_OF_MEMORY; }
        cudaMemcpyAsync(d_in, segs, n * 48, cudaMemcpyHostToDevice, t->stream);
    }
    expand_segments_kernel<<<grid_for(n, 256), 256, 0, t->stream>>>(d_in ? d_in : segs, n, t->d_soup);
    t->stats.kernel_launches += 1;
    cudaError_t e = cudaSuccess;
    if (ids) { e = cudaMalloc(&t->d_ids, n * 4); if (e == cudaSuccess) cudaMemcpyAsync(t->d_ids, ids, n * 4, kind, t->stream); }
    const cudaError_t e2 = cudaStreamSynchronize(t->stream);
    cudaFree(d_in);
    if (e != cudaSuccess || e2 != cudaSuccess || cudaGetLastError() != cudaSuccess) { free_input(t); t->err = "set_segments: CUDA failure"; return B200AABB_ERR_CUDA; }
    t->n = n; t->built = false;
    return B200AABB_OK;
}

int b200aabb_set_mesh(b200aabb_tree* t, const double* verts, uint64_t nv, const uint32_t* faces, const uint32_t* face_ids, uint64_t nf) { B2_LOCK(t);
    if (!t || (nf && (!verts || !faces))) return B200AABB_ERR_INVALID_ARGUMENT;
    for (uint64_t i = 0; i < 3 * nf; ++i) if (faces[i] >= nv) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "face references a vertex index >= nv");
    int rc = b200aabb_clear(t);
    if (rc) return rc;
    if (nf == 0) return B200AABB_OK;
    double* d_v = nullptr; uint32_t* d_f = nullptr;
    B2_CUDA(t, cudaMalloc(&d_v, nv * 24));
    cudaError_t e = cudaMalloc(&d_f, nf * 12);
    if (e != cudaSuccess) { cudaFree(d_v); t->err = "cudaMalloc(faces)"; return B200AABB_ERR_OUT_OF_MEMORY; }
    e = cudaMalloc(&t->d_soup, nf * 72 + kSoupSlack);
    if (e != cudaSuccess) { cudaFree(d_v); cudaFree(d_f); t->err = "cudaMalloc(soup)"; return B200AABB_ERR_OUT_OF_MEMORY; }
    cudaMemcpyAsync(d_v, verts, nv * 24, cudaMemcpyHostToDevice, t->stream);
    cudaMemcpyAsync(d_f, faces, nf * 12, cudaMemcpyHostToDevice, t->stream);
    cudaMemsetAsync(t->d_bounds, 0xFF, 3 * 8, t->stream);
    cudaMemsetAsync(t->d_bounds + 3, 0x00, 3 * 8, t->stream);
    gather_mesh_kernel<<<std::min<unsigned>(grid_for(nf, 256), 148 * 8), 256, 0, t->stream>>>(d_v, d_f, nf, t->d_soup, t->d_bounds);
    t->stats.kernel_launches += 1;
    if (face_ids) { e = cudaMalloc(&t->d_ids, nf * 4); if (e == cudaSuccess) cudaMemcpyAsync(t->d_ids, face_ids, nf * 4, cudaMemcpyHostToDevice, t->stream); }
    cudaError_t e2 = cudaStreamSynchronize(t->stream);
    cudaFree(d_v); cudaFree(d_f);
    if (e != cudaSuccess || e2 != cudaSuccess) { free_input(t); t->err = "set_mesh: CUDA failure"; return B200AABB_ERR_CUDA; }
    t->n = nf; t->built = false; t->bounds_valid = true;
    return B200AABB_OK;
}

// OFF text -> primitives on the device (ingest.cuh).  The host reads only the header (the "OFF" keyword, the three counts, comments before them); the body
// is tokenised and parsed by kernels.  Accepted: plain OFF (no per-vertex normals / colours / texture coordinates, no comments after the counts) whose faces are
// all triangles "3 i j k" — the files the tree of the north-star path is built from; anything else returns INVALID_ARGUMENT with the reason, and the caller
// falls back to its own reader + b200aabb_set_mesh.
__global__ void __launch_bounds__(256) gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
__global__ void __launch_bounds__(256) scatter_f64_kernel(const double* __restrict__ val, const uint32_t* __restrict__ idx, uint32_t n, double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[idx[i]] = val[i];
}
int b200aabb_set_mesh_off(b200aabb_tree* t, const char* text, uint64_t len) { B2_LOCK(t);
    if (!t || !text) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = b200aabb_clear(t);                              // (whatever happens below, the old primitives are gone: a failed ingest leaves an empty tree)
    if (rc) return rc;
    if (len >= (1ull << 32)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "set_mesh_off: at most 4 GiB of text");
    // ---- header (File_header_OFF.cpp: keyword, then n_vertices n_facets n_halfedges; '#' starts a comment)
    uint64_t pos = 0;
    auto skip_space_and_comments = [&]() {
        while (pos < len) {
            if (is_space(text[pos])) ++pos;
            else if (text[pos] == '#') { while (pos < len && text[pos] != '\n') ++pos; }
            else break;
        }
    };
    skip_space_and_comments();
    if (!(pos + 3 <= len && text[pos] == 'O' && text[pos + 1] == 'F' && text[pos + 2] == 'F' && (pos + 3 == len || is_space(text[pos + 3]) || text[pos + 3] == '#')))
        return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "set_mesh_off: the text does not start with the keyword OFF (COFF / NOFF / STOFF variants are not taken)");
    pos += 3;
    uint64_t counts[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) {
        skip_space_and_comments();
        uint64_t b = pos;
        while (pos < len && text[pos] >= '0' && text[pos] <= '9' && pos - b < 19) { counts[k] = counts[k] * 10 + (uint64_t)(text[pos] - '0'); ++pos; }
        if (pos == b || (pos < len && !is_space(text[pos]) && text[pos] != '#')) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "set_mesh_off: the header needs three counts (vertices, facets, edges)");
    }
    const uint64_t nv = counts[0], nf = counts[1];
    skip_space_and_comments();
    if (memchr(text + pos, '#', len - pos)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "set_mesh_off: comments after the header are not taken");
    if (nf >= (1ull << 30) || nv >= (1ull << 32)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "set_mesh_off: too many vertices or facets");
    if (nf == 0) return B200AABB_OK;
    B2_CUDA(t, cudaSetDevice(t->device));
    const uint64_t want_tokens = 3 * nv + 4 * nf;
    const uint64_t body = len - pos;
    const uint32_t ntiles = (uint32_t)((body + kTokTile - 1) / kTokTile);
    char* d_text = nullptr; uint32_t* d_tiles = nullptr; uint32_t* d_tok = nullptr; double* d_v = nullptr; uint32_t* d_f = nullptr; unsigned* d_flags = nullptr;
    uint32_t* d_und = nullptr; unsigned long long* d_total = nullptr;
    constexpr uint32_t kUndecidedCap = 1u << 16;
    auto release = [&]() { cudaFree(d_text); cudaFree(d_tiles); cudaFree(d_tok); cudaFree(d_v); cudaFree(d_f); cudaFree(d_flags); cudaFree(d_und); cudaFree(d_total); };
    cudaError_t e = cudaMalloc(&d_text, len);
    if (e == cudaSuccess) e = cudaMalloc(&d_tiles, std::max<uint64_t>(ntiles, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_tok, std::max<uint64_t>(want_tokens, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_v, std::max<uint64_t>(nv, 1) * 24);
    if (e == cudaSuccess) e = cudaMalloc(&d_f, nf * 12);
    if (e == cudaSuccess) e = cudaMalloc(&d_flags, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(&d_und, kUndecidedCap * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_total, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&t->d_soup, nf * 72 + kSoupSlack);
    if (e != cudaSuccess) { release(); free_input(t); t->err = "cudaMalloc(set_mesh_off)"; return B200AABB_ERR_OUT_OF_MEMORY; }
    cudaStream_t st = t->stream;
    static const bool timing = getenv("B200AABB_INGEST_TIMING") != nullptr;    // diagnostic: upload / kernels split on stderr
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (timing) { for (auto& x : ev) cudaEventCreate(&x); cudaEventRecord(ev[0], st); }
    cudaMemcpyAsync(d_text, text, len, cudaMemcpyHostToDevice, st);
    if (timing) cudaEventRecord(ev[1], st);
    cudaMemsetAsync(d_flags, 0, 2 * sizeof(unsigned), st);
    token_count_kernel<<<std::max(ntiles, 1u), kTokBlock, 0, st>>>(d_text, pos, len, d_tiles);
    token_scan_kernel<<<1, 1024, 0, st>>>(d_tiles, ntiles, d_total);
    token_offsets_kernel<<<std::max(ntiles, 1u), kTokBlock, 0, st>>>(d_text, pos, len, d_tiles, d_tok, want_tokens);
    unsigned long long total = 0;
    cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { release(); free_input(t); t->err = "set_mesh_off: CUDA failure (tokens)"; return B200AABB_ERR_CUDA; }
    if (total != want_tokens) {
        release(); free_input(t);
        return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "set_mesh_off: the body does not hold 3 numbers per vertex and 4 per facet (non-triangular facets, colours or normals?)");
    }
    if (nv) parse_vertices_kernel<<<grid_for(3 * nv, 256), 256, 0, st>>>(d_text, len, d_tok, 3 * nv, d_v, d_flags, d_und, kUndecidedCap);
    parse_faces_kernel<<<grid_for(nf, 256), 256, 0, st>>>(d_text, len, d_tok + 3 * nv, nf, nv, d_f, d_flags);
    unsigned flags[2] = {0, 0};
    cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    t->stats.kernel_launches += 5;
    if (e != cudaSuccess) { release(); free_input(t); t->err = "set_mesh_off: CUDA failure (parse)"; return B200AABB_ERR_CUDA; }
    if (flags[0] || flags[1] > kUndecidedCap) {
        release(); free_input(t);
        return fail(t, B200AABB_ERR_INVALID_ARGUMENT, (flags[0] & 1) ? "set_mesh_off: a vertex coordinate is not a decimal number" :
                    (flags[0] & 2) ? "set_mesh_off: a facet is not a triangle" : (flags[0] & 4) ? "set_mesh_off: a facet references a vertex index >= the vertex count" :
                    "set_mesh_off: too many literals of more than 19 digits");
    }
    if (flags[1]) {   // literals of more than 19 significant digits whose dropped digits decide the rounding: strtod on the caller's text
        std::vector<uint32_t> und(flags[1]), off(flags[1]);
        std::vector<double> val(flags[1]);
        uint32_t* d_off = nullptr; double* d_val = nullptr;
        if (cudaMalloc(&d_off, flags[1] * 4) != cudaSuccess || cudaMalloc(&d_val, flags[1] * 8) != cudaSuccess) {
            cudaFree(d_off); release(); free_input(t); t->err = "cudaMalloc(set_mesh_off)"; return B200AABB_ERR_OUT_OF_MEMORY;
        }
        gather_u32_kernel<<<grid_for(flags[1], 256), 256, 0, st>>>(d_tok, d_und, flags[1], d_off);
        cudaMemcpyAsync(und.data(), d_und, flags[1] * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(off.data(), d_off, flags[1] * 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        for (unsigned i = 0; i < flags[1]; ++i) {
            uint64_t b = off[i], en = b;
            while (en < len && !is_space(text[en])) ++en;
            val[i] = strtod(std::string(text + b, text + en).c_str(), nullptr);
        }
        cudaMemcpyAsync(d_val, val.data(), flags[1] * 8, cudaMemcpyHostToDevice, st);
        scatter_f64_kernel<<<grid_for(flags[1], 256), 256, 0, st>>>(d_val, d_und, flags[1], d_v);
        cudaStreamSynchronize(st);
        cudaFree(d_off); cudaFree(d_val);
    }
    cudaMemsetAsync(t->d_bounds, 0xFF, 3 * 8, st);
    cudaMemsetAsync(t->d_bounds + 3, 0x00, 3 * 8, st);
    gather_mesh_kernel<<<std::min<unsigned>(grid_for(nf, 256), 148 * 8), 256, 0, st>>>(d_v, d_f, nf, t->d_soup, t->d_bounds);
    t->stats.kernel_launches += 1;
    if (timing) cudaEventRecord(ev[2], st);
    e = cudaStreamSynchronize(st);
    if (timing) {
        float up = 0, run = 0;
        cudaEventElapsedTime(&up, ev[0], ev[1]); cudaEventElapsedTime(&run, ev[1], ev[2]);
        fprintf(stderr, "[b200aabb ingest] %.1f MB of text: upload %.2f ms, tokenise + parse + gather %.2f ms (%llu tokens)\n", len / 1e6, up, run, (unsigned long long)want_tokens);
        for (auto& x : ev) cudaEventDestroy(x);
    }
    release();
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { free_input(t); t->err = "set_mesh_off: CUDA failure (gather)"; return B200AABB_ERR_CUDA; }
    t->n = nf; t->built = false; t->bounds_valid = true;
    return B200AABB_OK;
}

// bottom-up boxes over the existing topology from the current leaf records (shared by b200aabb_refit)
static int refit_boxes(b200aabb_tree* t, cudaStream_t st) {
    const uint64_t n = t->n;
    if (n < 2) return B200AABB_OK;
    const uint64_t refit_blocks = (n + kRefitBlock - 1) / kRefitBlock, frontier_cap = refit_blocks * kRefitFrontierPerBlock;
    { const int rc_ = ensure_build_scratch(t, n); if (rc_) return rc_; }
    RefitEntry* const frontier = t->bs.frontier; unsigned* const frontier_count = t->bs.frontier_count;
    cudaMemsetAsync(frontier_count, 0, sizeof(unsigned), st);
    refit_local_kernel<false><<<(unsigned)refit_blocks, kRefitBlock, 0, st>>>(t->d_tris, n, t->d_nodes, t->d_leaf_parent, frontier, frontier_count, (unsigned)frontier_cap, EmitArgs{});
    refit_global_kernel<<<(unsigned)((frontier_cap + 255) / 256), 256, 0, st>>>(t->d_nodes, t->d_leaf_parent, frontier, frontier_count, (unsigned)frontier_cap);
    t->stats.kernel_launches += 2;
    return cudaGetLastError() == cudaSuccess ? B200AABB_OK : fail(t, B200AABB_ERR_CUDA, "refit: CUDA failure");
}

static int refit_from_soup(b200aabb_tree* t) {
    cudaStream_t st = t->stream;
    const uint64_t n = t->n;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    B2_CUDA(t, cudaEventRecord(t->ev_a, st));
    if (!t->bounds_valid) {
        cudaMemsetAsync(t->d_bounds, 0xFF, 3 * 8, st);
        cudaMemsetAsync(t->d_bounds + 3, 0x00, 3 * 8, st);
        scene_bounds_kernel<<<std::min<unsigned>(blocks, 148 * 8), 256, 0, st>>>(t->d_soup, n, t->d_bounds);
        t->bounds_valid = true;
        t->stats.kernel_launches += 1;
    }
    refresh_leaves_kernel<<<blocks, 256, 0, st>>>(t->d_soup, n, t->d_tris, t->d_trisf);
    t->stats.kernel_launches += 1;
    int rc = refit_boxes(t, st);
    if (rc) return rc;
    root_box_kernel<<<1, 32, 0, st>>>(t->d_bounds, t->d_rootbox);
    B2_CUDA(t, cudaEventRecord(t->ev_b, st));
    B2_CUDA(t, cudaMemcpyAsync(t->h_bbox, t->d_rootbox, 48, cudaMemcpyDeviceToHost, st));
    B2_CUDA(t, cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, t->ev_a, t->ev_b);
    t->stats.last_build_ms = ms;
    t->mag = 0.0;
    for (int a = 0; a < 6; ++a) t->mag = std::max(t->mag, fabs(t->h_bbox[a]));
    return B200AABB_OK;
}

static int refit_precheck(b200aabb_tree* t, uint64_t n) {
    if (t->imported || !t->d_soup) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "refit needs a tree built from inserted primitives");
    if (n != t->n) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "refit keeps the topology: the number of primitives must not change (use set_* + build)");
    int rc = ensure_built(t);
    if (rc) return rc;
    { const int rc_ = sync_pipeline(t); if (rc_) return rc_; }
    return B200AABB_OK;
}

int b200aabb_refit(b200aabb_tree* t, const double* tris, uint64_t n, b200aabb_memspace space) {
    if (!t || (n && !tris)) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_LOCK(t);
    B2_CUDA(t, cudaSetDevice(t->device));
    { const int rc_ = need_triangles(t, "refit"); if (rc_) return rc_; }
    int rc = refit_precheck(t, n);
    if (rc || n == 0) return rc;
    B2_CUDA(t, cudaMemcpyAsync(t->d_soup, tris, n * 72, space == B200AABB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, t->stream));
    t->bounds_valid = false;
    return refit_from_soup(t);
}

int b200aabb_refit_mesh(b200aabb_tree* t, const double* verts, uint64_t nv, const uint32_t* faces, uint64_t nf) {
    if (!t || (nf && (!verts || !faces))) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_LOCK(t);
    B2_CUDA(t, cudaSetDevice(t->device));
    for (uint64_t i = 0; i < 3 * nf; ++i) if (faces[i] >= nv) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "face references a vertex index >= nv");
    { const int rc_ = need_triangles(t, "refit_mesh"); if (rc_) return rc_; }
    int rc = refit_precheck(t, nf);
    if (rc || nf == 0) return rc;
    double* d_v = nullptr; uint32_t* d_f = nullptr;
    if (cudaMalloc(&d_v, nv * 24) != cudaSuccess || cudaMalloc(&d_f, nf * 12) != cudaSuccess) { cudaFree(d_v); cudaFree(d_f); return fail(t, B200AABB_ERR_OUT_OF_MEMORY, "cudaMalloc(refit_mesh)"); }
    cudaMemcpyAsync(d_v, verts, nv * 24, cudaMemcpyHostToDevice, t->stream);
    cudaMemcpyAsync(d_f, faces, nf * 12, cudaMemcpyHostToDevice, t->stream);
    cudaMemsetAsync(t->d_bounds, 0xFF, 3 * 8, t->stream);
    cudaMemsetAsync(t->d_bounds + 3, 0x00, 3 * 8, t->stream);
    gather_mesh_kernel<<<std::min<unsigned>(grid_for(nf, 256), 148 * 8), 256, 0, t->stream>>>(d_v, d_f, nf, t->d_soup, t->d_bounds);
    t->stats.kernel_launches += 1;
    t->bounds_valid = true;
    rc = refit_from_soup(t);
    cudaFree(d_v); cudaFree(d_f);
    return rc;
}

int b200aabb_build(b200aabb_tree* t) { B2_LOCK(t); if (!t) return B200AABB_ERR_INVALID_ARGUMENT; return do_build(t); }
int b200aabb_accelerate_distance_queries(b200aabb_tree* t) { B2_LOCK(t); if (!t) return B200AABB_ERR_INVALID_ARGUMENT; return ensure_built(t); }
int b200aabb_size(const b200aabb_tree* t, uint64_t* n) { B2_LOCK(t); if (!t || !n) return B200AABB_ERR_INVALID_ARGUMENT; *n = t->n; return B200AABB_OK; }
int b200aabb_bbox(b200aabb_tree* t, double out[6]) { B2_LOCK(t);
    if (!t || !out) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = ensure_built(t);
    if (rc) return rc;
    memcpy(out, t->h_bbox, 48);
    return B200AABB_OK;
}

int b200aabb_do_intersect(b200aabb_tree* t, const double* q, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint8_t* out) { B2_LOCK(t);
    if (!t || (n && (!q || !out)) || !valid_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (!kind_fits_primitives(t, kind)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "a tree of Segment_3 primitives takes Plane_3 and Triangle_3 queries");
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) {   // empty tree: false (aabb_test_empty_tree.cpp:44-50)
        if (space == B200AABB_DEVICE) { B2_CUDA(t, cudaMemsetAsync(out, 0, n, t->stream)); } else memset(out, 0, n);
        return B200AABB_OK;
    }
    return run_batch(t, q, query_bytes(kind), n, space, query_shape(kind), {{out, 1, 0}}, [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
        launch_intersect<kAnyHit>(t, kind, st, (const double*)in, cnt, perm, (uint8_t*)o[0], nullptr, nullptr, nullptr);
    });
}

int b200aabb_number_of_intersected_primitives(b200aabb_tree* t, const double* q, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint32_t* out) { B2_LOCK(t);
    if (!t || (n && (!q || !out)) || !valid_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (!kind_fits_primitives(t, kind)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "a tree of Segment_3 primitives takes Plane_3 and Triangle_3 queries");
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) {
        if (space == B200AABB_DEVICE) { B2_CUDA(t, cudaMemsetAsync(out, 0, n * 4, t->stream)); } else memset(out, 0, n * 4);
        return B200AABB_OK;
    }
    return run_batch(t, q, query_bytes(kind), n, space, query_shape(kind), {{out, 4, 0}}, [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
        launch_intersect<kCount>(t, kind, st, (const double*)in, cnt, perm, nullptr, (uint32_t*)o[0], nullptr, nullptr);
    });
}

int b200aabb_any_intersected_primitive(b200aabb_tree* t, const double* q, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint32_t* out) { B2_LOCK(t);
    if (!t || (n && (!q || !out)) || !valid_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (!kind_fits_primitives(t, kind)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "a tree of Segment_3 primitives takes Plane_3 and Triangle_3 queries");
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) {   // empty tree: nullopt
        if (space == B200AABB_DEVICE) { B2_CUDA(t, cudaMemsetAsync(out, 0xFF, n * 4, t->stream)); } else memset(out, 0xFF, n * 4);
        return B200AABB_OK;
    }
    return run_batch(t, q, query_bytes(kind), n, space, query_shape(kind), {{out, 4, 0}}, [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
        launch_intersect<kAnyId>(t, kind, st, (const double*)in, cnt, perm, nullptr, (uint32_t*)o[0], nullptr, nullptr);
    });
}

int b200aabb_all_intersected_primitives(b200aabb_tree* t, const double* q, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space,
                                        const uint64_t* offsets, uint32_t* ids) { B2_LOCK(t);
    if (!t || (n && (!q || !offsets)) || !valid_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (!kind_fits_primitives(t, kind)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "a tree of Segment_3 primitives takes Plane_3 and Triangle_3 queries");
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    B2_CUDA(t, cudaSetDevice(t->device));
    const size_t qbytes = query_bytes(kind);
    // one launch over the whole batch, in the caller's order (the id lists are variable-length: no chunked staging of the output)
    const double* d_q = q; const uint64_t* d_off = offsets; uint32_t* d_ids = ids;
    uint64_t total = 0;
    void* tmp[3] = {nullptr, nullptr, nullptr};
    auto release = [&]() { for (void* p : tmp) cudaFree(p); };
    if (space == B200AABB_HOST) {
        if (offsets[0] != 0) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets[0] must be 0");
        for (uint64_t i = 0; i < n; ++i) if (offsets[i + 1] < offsets[i]) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
        total = offsets[n];
        if (total && !ids) return B200AABB_ERR_INVALID_ARGUMENT;
        if (t->n == 0) return total == 0 ? B200AABB_OK : fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets disagree with the number of intersected primitives");
        cudaError_t e = cudaMalloc(&tmp[0], n * qbytes);
        if (e == cudaSuccess) e = cudaMalloc(&tmp[1], (n + 1) * 8);
        if (e == cudaSuccess) e = cudaMalloc(&tmp[2], std::max<uint64_t>(total, 1) * 4);
        if (e != cudaSuccess) { release(); t->err = "cudaMalloc(all_intersected_primitives staging)"; return B200AABB_ERR_OUT_OF_MEMORY; }
        cudaMemcpyAsync(tmp[0], q, n * qbytes, cudaMemcpyHostToDevice, t->stream);
        cudaMemcpyAsync(tmp[1], offsets, (n + 1) * 8, cudaMemcpyHostToDevice, t->stream);
        d_q = (const double*)tmp[0]; d_off = (const uint64_t*)tmp[1]; d_ids = (uint32_t*)tmp[2];
    } else if (t->n == 0) return B200AABB_OK;
    cudaMemsetAsync(t->d_counters, 0, 8 * sizeof(unsigned long long), t->stream);
    cudaEventRecord(t->ev_a, t->stream);
    const uint32_t* perm = nullptr;
    rc = coherence_order(t, kDeviceScratch, t->stream, d_q, n, query_shape(kind), &perm);
    if (rc) { release(); return rc; }
    launch_intersect<kList>(t, kind, t->stream, d_q, n, perm, nullptr, nullptr, d_off, d_ids);
    t->stats.kernel_launches += 1;
    cudaEventRecord(t->ev_b, t->stream);
    t->query_timing_pending = true;
    if (space == B200AABB_HOST) {
        if (total) cudaMemcpyAsync(ids, d_ids, total * 4, cudaMemcpyDeviceToHost, t->stream);
        unsigned long long c2 = 0;
        cudaMemcpyAsync(&c2, t->d_counters + 2, sizeof(c2), cudaMemcpyDeviceToHost, t->stream);
        const cudaError_t e = cudaStreamSynchronize(t->stream);
        release();
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { t->err = "all_intersected_primitives: CUDA failure"; return B200AABB_ERR_CUDA; }
        if (c2 & 4ull) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets disagree with the number of intersected primitives");
        if (c2) return fail(t, B200AABB_ERR_INCOMPLETE, "traversal stack overflow: results incomplete");
    }
    B2_CUDA(t, cudaGetLastError());
    return B200AABB_OK;
}

int b200aabb_do_intersect_tree(b200aabb_tree* t, b200aabb_tree* other, int* result, uint32_t* prim, uint32_t* other_prim) {
    return b200aabb_do_intersect_tree_transformed(t, nullptr, other, nullptr, result, prim, other_prim);
}
int b200aabb_do_intersect_tree_transformed(b200aabb_tree* t, const double* xf, b200aabb_tree* other, const double* other_xf, int* result, uint32_t* prim,
                                           uint32_t* other_prim) {
    if (!t || !other || !result) return B200AABB_ERR_INVALID_ARGUMENT;
    // both handles are locked, in address order (two threads testing (a, b) and (b, a) must not dead-lock)
    std::unique_lock<std::recursive_mutex> l1((t < other ? t : other)->mu), l2;
    if (t != other) l2 = std::unique_lock<std::recursive_mutex>((t < other ? other : t)->mu);
    *result = 0;
    if (prim) *prim = kNone;
    if (other_prim) *other_prim = kNone;
    if (t->device != other->device) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "do_intersect(tree): both trees must live on the same device");
    if (t->segments || other->segments) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "do_intersect(tree): not offered on a tree of Segment_3 primitives");
    int rc = ensure_built(t);
    if (rc) return rc;
    rc = ensure_built(other);
    if (rc) { t->err = other->err; return rc; }
    if (t->n == 0 || other->n == 0) return B200AABB_OK;                 // an empty tree intersects nothing (AABB_tree.h:516-525)
    {   // the first node test of the reference's traversal: when the (images of the) root boxes are disjoint nothing can intersect
        double ba[6], bb[6];
        if (!image_of_box(xf, t->h_bbox, ba) || !image_of_box(other_xf, other->h_bbox, bb)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "do_intersect(tree): a transformation is not finite");
        if (ba[0] > bb[3] || bb[0] > ba[3] || ba[1] > bb[4] || bb[1] > ba[4] || ba[2] > bb[5] || bb[2] > ba[5]) return B200AABB_OK;
    }
    B2_CUDA(t, cudaSetDevice(t->device));
    rc = drain(other);                                                   // (its build, or batches still in flight, run on other streams)
    if (rc) { t->err = other->err; return rc; }
    // the queries are the primitives of the SMALLER tree, the traversed tree is the larger one: the answer is symmetric
    // (Do_intersect_3(Triangle_3, Triangle_3) decides exactly), the work is (number of queries) x log(size of the traversed tree)
    const bool t_is_big = t->n >= other->n;                               // (not decided by comparing handles: a tree may be tested against itself under another transformation)
    b200aabb_tree* big = t_is_big ? t : other;
    b200aabb_tree* small = t_is_big ? other : t;
    unsigned* const d_w = t->d_witness;                                  // (kept with the handle: no allocation per test)
    const unsigned init[3] = {0u, kNone, kNone};
    cudaMemcpyAsync(d_w, init, sizeof(init), cudaMemcpyHostToDevice, t->stream);
    rc = reset_counters(big, t->stream);
    if (rc) return rc;
    cudaEventRecord(t->ev_a, t->stream);
    XfPair x;
    static const double kIdentity[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const double* xf_small = t_is_big ? other_xf : xf;
    const double* xf_big = t_is_big ? xf : other_xf;
    memcpy(x.q.m, xf_small ? xf_small : kIdentity, sizeof(x.q.m));
    memcpy(x.b.m, xf_big ? xf_big : kIdentity, sizeof(x.b.m));
    if (xf || other_xf) tree_tree_kernel<true><<<grid_for(small->n, 128), 128, 0, t->stream>>>(tree_ref(big, nullptr), small->d_tris, small->n, d_w, x);
    else tree_tree_kernel<false><<<grid_for(small->n, 128), 128, 0, t->stream>>>(tree_ref(big, nullptr), small->d_tris, small->n, d_w, x);
    t->stats.kernel_launches += 1;
    cudaEventRecord(t->ev_b, t->stream);
    t->query_timing_pending = true;
    unsigned w[3] = {0, 0, 0};
    unsigned long long c2 = 0;
    cudaMemcpyAsync(w, d_w, sizeof(w), cudaMemcpyDeviceToHost, t->stream);
    cudaMemcpyAsync(&c2, big->d_counters + 2, sizeof(c2), cudaMemcpyDeviceToHost, t->stream);
    const cudaError_t e = cudaStreamSynchronize(t->stream);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { t->err = "do_intersect(tree): CUDA failure"; return B200AABB_ERR_CUDA; }
    if (c2 && !w[0]) return fail(t, B200AABB_ERR_INCOMPLETE, "traversal stack overflow: results incomplete");
    *result = w[0] ? 1 : 0;
    if (w[0]) {                                                          // w[1]: id in `small`, w[2]: id in `big`
        if (prim) *prim = t_is_big ? w[2] : w[1];
        if (other_prim) *other_prim = t_is_big ? w[1] : w[2];
    }
    return B200AABB_OK;
}

int b200aabb_number_of_intersections(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, uint32_t* out) {
    return b200aabb_number_of_intersections_q(t, rays, n, B200AABB_RAY, space, out);
}
int b200aabb_number_of_intersections_q(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint32_t* out) { B2_LOCK(t);
    if (!t || (n && (!rays || !out)) || !valid_object_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    { const int rc_ = need_triangles(t, "intersection objects"); if (rc_) return rc_; }
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) {
        if (space == B200AABB_DEVICE) { B2_CUDA(t, cudaMemsetAsync(out, 0, n * 4, t->stream)); } else memset(out, 0, n * 4);
        return B200AABB_OK;
    }
    return run_batch(t, rays, query_bytes(kind), n, space, query_shape(kind), {{out, 4, 0}}, [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
        launch_objects<kCount>(t, kind, st, (const double*)in, cnt, perm, (uint32_t*)o[0], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    });
}

int b200aabb_any_intersection(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, uint32_t* prim, double* pt, double* pt2, uint8_t* flags) {
    return b200aabb_any_intersection_q(t, rays, n, B200AABB_RAY, space, prim, pt, pt2, flags);
}
int b200aabb_any_intersection_q(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint32_t* prim, double* pt, double* pt2,
                                uint8_t* flags) { B2_LOCK(t);
    if (!t || (n && !rays) || !valid_object_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    { const int rc_ = need_triangles(t, "intersection objects"); if (rc_) return rc_; }
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) return b200aabb_first_intersection(t, rays, n, space, prim, nullptr, pt, pt2, flags);   // nullopt for every ray
    return run_batch(t, rays, query_bytes(kind), n, space, query_shape(kind), {{prim, 4, 0}, {pt, 24, 0}, {pt2, 24, 0}, {flags, 1, 0}},
                     [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
                         launch_objects<kAnyId>(t, kind, st, (const double*)in, cnt, perm, nullptr, nullptr, nullptr, nullptr, nullptr, (uint32_t*)o[0], (uint8_t*)o[3], (double*)o[1], (double*)o[2]);
                     });
}

int b200aabb_all_intersections(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, const uint64_t* offsets, uint32_t* ids,
                               uint8_t* kinds, double* points) {
    return b200aabb_all_intersections_q(t, rays, n, B200AABB_RAY, space, offsets, ids, kinds, points);
}
int b200aabb_all_intersections_q(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, const uint64_t* offsets, uint32_t* ids,
                                 uint8_t* kinds, double* points) { B2_LOCK(t);
    if (!t || (n && (!rays || !offsets)) || !valid_object_kind(kind)) return B200AABB_ERR_INVALID_ARGUMENT;
    { const int rc_ = need_triangles(t, "intersection objects"); if (rc_) return rc_; }
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    B2_CUDA(t, cudaSetDevice(t->device));
    const double* d_q = rays; const uint64_t* d_off = offsets; uint32_t* d_ids = ids; uint8_t* d_kinds = kinds; double* d_pts = points;
    uint64_t total = 0;
    void* tmp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    auto release = [&]() { for (void* p : tmp) cudaFree(p); };
    if (space == B200AABB_HOST) {
        if (offsets[0] != 0) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets[0] must be 0");
        for (uint64_t i = 0; i < n; ++i) if (offsets[i + 1] < offsets[i]) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
        total = offsets[n];
        if (total && (!ids || !kinds || !points)) return B200AABB_ERR_INVALID_ARGUMENT;
        if (t->n == 0) return total == 0 ? B200AABB_OK : fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets disagree with the number of intersections");
        const uint64_t room = std::max<uint64_t>(total, 1);
        cudaError_t e = cudaMalloc(&tmp[0], n * query_bytes(kind));
        if (e == cudaSuccess) e = cudaMalloc(&tmp[1], (n + 1) * 8);
        if (e == cudaSuccess) e = cudaMalloc(&tmp[2], room * 4);
        if (e == cudaSuccess) e = cudaMalloc(&tmp[3], room);
        if (e == cudaSuccess) e = cudaMalloc(&tmp[4], room * 48);
        if (e != cudaSuccess) { release(); t->err = "cudaMalloc(all_intersections staging)"; return B200AABB_ERR_OUT_OF_MEMORY; }
        cudaMemcpyAsync(tmp[0], rays, n * query_bytes(kind), cudaMemcpyHostToDevice, t->stream);
        cudaMemcpyAsync(tmp[1], offsets, (n + 1) * 8, cudaMemcpyHostToDevice, t->stream);
        d_q = (const double*)tmp[0]; d_off = (const uint64_t*)tmp[1]; d_ids = (uint32_t*)tmp[2]; d_kinds = (uint8_t*)tmp[3]; d_pts = (double*)tmp[4];
    } else if (t->n == 0) return B200AABB_OK;
    cudaMemsetAsync(t->d_counters, 0, 8 * sizeof(unsigned long long), t->stream);
    cudaEventRecord(t->ev_a, t->stream);
    const uint32_t* perm = nullptr;
    rc = coherence_order(t, kDeviceScratch, t->stream, d_q, n, query_shape(kind), &perm);
    if (rc) { release(); return rc; }
    launch_objects<kList>(t, kind, t->stream, d_q, n, perm, nullptr, d_off, d_ids, d_kinds, d_pts, nullptr, nullptr, nullptr, nullptr);
    t->stats.kernel_launches += 1;
    cudaEventRecord(t->ev_b, t->stream);
    t->query_timing_pending = true;
    if (space == B200AABB_HOST) {
        if (total) {
            cudaMemcpyAsync(ids, d_ids, total * 4, cudaMemcpyDeviceToHost, t->stream);
            cudaMemcpyAsync(kinds, d_kinds, total, cudaMemcpyDeviceToHost, t->stream);
            cudaMemcpyAsync(points, d_pts, total * 48, cudaMemcpyDeviceToHost, t->stream);
        }
        unsigned long long c2 = 0;
        cudaMemcpyAsync(&c2, t->d_counters + 2, sizeof(c2), cudaMemcpyDeviceToHost, t->stream);
        const cudaError_t e = cudaStreamSynchronize(t->stream);
        release();
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { t->err = "all_intersections: CUDA failure"; return B200AABB_ERR_CUDA; }
        if (c2 & 4ull) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "offsets disagree with the number of intersections");
        if (c2) return fail(t, B200AABB_ERR_INCOMPLETE, "traversal stack overflow: results incomplete");
    }
    B2_CUDA(t, cudaGetLastError());
    return B200AABB_OK;
}

static int first_intersection_impl(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, const uint32_t* skip_ids, const uint32_t* skip_mask,
                                   uint32_t mask_words, uint32_t* prim, double* tpar, double* pt, double* pt2, uint8_t* flags) {
    if (!t || (n && !rays) || (skip_mask == nullptr && mask_words != 0)) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_LOCK(t);
    { const int rc_ = need_triangles(t, "first_intersection"); if (rc_) return rc_; }
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) {   // nullopt for every ray
        if (space == B200AABB_DEVICE) {
            if (prim) B2_CUDA(t, cudaMemsetAsync(prim, 0xFF, n * 4, t->stream));
            if (tpar) B2_CUDA(t, cudaMemsetAsync(tpar, 0, n * 8, t->stream));
            if (pt) B2_CUDA(t, cudaMemsetAsync(pt, 0, n * 24, t->stream));
            if (pt2) B2_CUDA(t, cudaMemsetAsync(pt2, 0, n * 24, t->stream));
            if (flags) B2_CUDA(t, cudaMemsetAsync(flags, 0, n, t->stream));
        } else {
            if (prim) memset(prim, 0xFF, n * 4);
            if (tpar) memset(tpar, 0, n * 8);
            if (pt) memset(pt, 0, n * 24);
            if (pt2) memset(pt2, 0, n * 24);
            if (flags) memset(flags, 0, n);
        }
        return B200AABB_OK;
    }
    // HOST batches: the mask travels once (before the chunks), the per-ray skip ids chunk by chunk next to the rays
    uint32_t* d_mask = nullptr;
    if (space == B200AABB_HOST) {
        B2_CUDA(t, cudaSetDevice(t->device));
        if (skip_mask && mask_words) {
            if (cudaMalloc(&d_mask, (size_t)mask_words * 4) != cudaSuccess) return fail(t, B200AABB_ERR_OUT_OF_MEMORY, "cudaMalloc(skip mask)");
            cudaMemcpy(d_mask, skip_mask, (size_t)mask_words * 4, cudaMemcpyHostToDevice);
        }
        if (skip_ids) {
            const size_t need = 4 * (size_t)kChunkMax;
            if (t->stage_aux_bytes < need) {
                for (int s = 0; s < kPipeSlots; ++s) { cudaFree(t->d_stage_aux[s]); t->d_stage_aux[s] = nullptr; }
                for (int s = 0; s < kPipeSlots; ++s) if (cudaMalloc(&t->d_stage_aux[s], need) != cudaSuccess) { cudaFree(d_mask); return fail(t, B200AABB_ERR_OUT_OF_MEMORY, "cudaMalloc(skip ids)"); }
                t->stage_aux_bytes = need;
            }
        }
    }
    const uint32_t* mask_dev = space == B200AABB_HOST ? d_mask : skip_mask;
    rc = run_batch(t, rays, 48, n, space, kLineQuery, {{prim, 4, 0}, {tpar, 8, 0}, {pt, 24, 0}, {pt2, 24, 0}, {flags, 1, 0}},
                   [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
                       const uint32_t* sk = skip_ids;
                       if (skip_ids && space == B200AABB_HOST) {
                           cudaMemcpyAsync(t->d_stage_aux[t->cur_slot], skip_ids + t->cur_offset, cnt * 4, cudaMemcpyHostToDevice, st);
                           sk = t->d_stage_aux[t->cur_slot];
                       }
                       if (sk || mask_dev)
                           first_intersection_kernel<true><<<grid_for(cnt, 128), 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, (uint32_t*)o[0], (double*)o[1],
                                                                                                (double*)o[2], (double*)o[3], (uint8_t*)o[4], sk, mask_dev, mask_words);
                       else if (t->opt_ray_kernel == 1 && t->n >= 2 && t->d_ray_cursor) {
                           unsigned long long* cursor = t->d_ray_cursor + (space == B200AABB_HOST ? kHostScratch : kDeviceScratch);
                           cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st);
                           const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)t->ray_grid, (cnt + 127) / 128);
                           first_intersection_stream_kernel<<<grid, 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, cursor, (uint32_t*)o[0], (double*)o[1], (double*)o[2],
                                                                                  (double*)o[3], (uint8_t*)o[4]);
                       } else
                           first_intersection_kernel<false><<<grid_for(cnt, 128), 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, (uint32_t*)o[0], (double*)o[1],
                                                                                                 (double*)o[2], (double*)o[3], (uint8_t*)o[4], nullptr, nullptr, 0);
                   });
    if (d_mask) { if (t->async_pending) drain(t); cudaFree(d_mask); }
    return rc;
}

int b200aabb_first_intersection(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, uint32_t* prim, double* tpar, double* pt,
                                double* pt2, uint8_t* flags) {
    return first_intersection_impl(t, rays, n, space, nullptr, nullptr, 0, prim, tpar, pt, pt2, flags);
}

int b200aabb_first_intersection_skip(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, const uint32_t* skip_ids, const uint32_t* skip_mask,
                                     uint32_t mask_words, uint32_t* prim, double* tpar, double* pt, double* pt2, uint8_t* flags) {
    return first_intersection_impl(t, rays, n, space, skip_ids, skip_mask, mask_words, prim, tpar, pt, pt2, flags);
}

int b200aabb_closest_point(b200aabb_tree* t, const double* pts, uint64_t n, b200aabb_memspace space, double* cp, uint32_t* prim, double* sqd, uint8_t* flags) { B2_LOCK(t);
    if (!t || (n && !pts)) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = ensure_built(t);
    if (rc) return rc;
    if (t->n == 0) return fail(t, B200AABB_ERR_EMPTY_TREE, "distance query on an empty tree (CGAL_precondition(!empty()))");
    if (n == 0) return B200AABB_OK;
    // the warp-cooperative kernel packs (query slot, node / leaf) in 32 bits and needs an internal root
    // (Segment_3 primitives: the pool kernel, which takes them by a flag, or one thread per query; kernels 2 and 3 are specialised on triangles)
    const bool warp_kernel = t->opt_closest_kernel >= 1 && t->n >= 2 && t->n <= kWpMaxPrims && !(t->segments && t->opt_closest_kernel != 1);
    if (warp_kernel) { for (int s = 0; s < kScratchSlots; ++s) { rc = ensure_warp(t, s); if (rc) return rc; } }
    const bool two_pass = warp_kernel && t->opt_closest_kernel == 3 && n < (1ull << 32);
    if (two_pass) { rc = ensure_filter(t, space == B200AABB_HOST ? kHostScratch : kDeviceScratch, space == B200AABB_HOST ? std::min<uint64_t>(n, kPointChunkMax) : n); if (rc) return rc; }
    return run_batch(t, pts, 24, n, space, kPointQuery, {{cp, 24, 0}, {prim, 4, 0}, {sqd, 8, 0}, {flags, 1, 0}},
                     [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
                         if (!warp_kernel || cnt >= (1ull << 32)) {
                             if (t->segments) closest_point_kernel<true><<<grid_for(cnt, 128), 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, (double*)o[0], (uint32_t*)o[1], (double*)o[2], (uint8_t*)o[3]);
                             else closest_point_kernel<false><<<grid_for(cnt, 128), 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, (double*)o[0], (uint32_t*)o[1], (double*)o[2], (uint8_t*)o[3]);
                             return;
                         }
                         const int slot = space == B200AABB_HOST ? kHostScratch : kDeviceScratch;
                         cudaMemsetAsync(t->warp[slot].cursor, 0, sizeof(unsigned long long), st);
                         if (two_pass && cnt <= t->warp[slot].fc_cap) {
                             // pass 1 (binary32 filter) -> pass 2 (exact resolve) -> the pool kernel for the queries whose candidate list overflowed
                             auto& w = t->warp[slot];
                             FilterClosestArgs f;
                             f.nodes = t->d_nodes; f.tris = t->d_tris; f.trisf = t->d_trisf; f.root = t->root; f.mag = t->mag; f.perm = perm; f.counters = t->d_counters;
                             f.pts = (const double*)in; f.n = cnt; f.cand = w.fc_cand; f.ncand = w.fc_ncand; f.overflow = w.fc_overflow; f.overflow_count = w.fc_count;
                             f.cp = (double*)o[0]; f.prim = (uint32_t*)o[1]; f.sqd = (double*)o[2]; f.flags = (uint8_t*)o[3];
                             f.flag_tie = B200AABB_FLAG_TIE; f.flag_exact = B200AABB_FLAG_EXACT;
                             cudaMemsetAsync(w.fc_count, 0, sizeof(unsigned long long), st);
                             closest_filter_kernel<<<grid_for(cnt, 128), 128, 0, st>>>(f);
                             closest_resolve_kernel<<<grid_for(cnt, 128), 128, 0, st>>>(f);
                             WarpClosestArgs a;
                             a.nodes = t->d_nodes; a.tris = t->d_tris; a.trisf = t->d_trisf; a.mag = t->mag; a.perm = w.fc_overflow; a.counters = t->d_counters;
                             a.pts = (const double*)in; a.n = cnt; a.n_dev = w.fc_count; a.cp = f.cp; a.prim = f.prim; a.sqd = f.sqd; a.flags = f.flags;
                             a.cursor = w.cursor; a.spill = w.spill; a.slots64 = w.slots64;
                             a.exact_go = t->opt_warp_exact_go > 0 ? t->opt_warp_exact_go : 160;
                             a.flag_tie = B200AABB_FLAG_TIE; a.flag_exact = B200AABB_FLAG_EXACT;
                             closest_point_warp_kernel<false><<<(unsigned)std::min<uint64_t>((uint64_t)t->warp_grid, (cnt + 31) / 32), 32, 0, st>>>(a);
                             t->stats.kernel_launches += 2;
                             return;
                         }
                         if (t->opt_closest_kernel == 2) {
                             PairClosestArgs a;
                             a.nodes = t->d_nodes; a.tris = t->d_tris; a.trisf = t->d_trisf; a.mag = t->mag; a.perm = perm; a.counters = t->d_counters;
                             a.pts = (const double*)in; a.n = cnt; a.cp = (double*)o[0]; a.prim = (uint32_t*)o[1]; a.sqd = (double*)o[2]; a.flags = (uint8_t*)o[3];
                             a.cursor = t->warp[slot].cursor; a.spill = t->warp[slot].spill; a.slots64 = t->warp[slot].slots64;
                             a.exact_go = t->opt_warp_exact_go > 0 ? std::min(t->opt_warp_exact_go, kPpExactCap - 32) : 64;
                             a.flag_tie = B200AABB_FLAG_TIE; a.flag_exact = B200AABB_FLAG_EXACT;
                             const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)t->pair_grid, (cnt + kPpChunk - 1) / kPpChunk);
                             closest_point_pair_kernel<<<grid, 32, 0, st>>>(a);
                             return;
                         }
                         WarpClosestArgs a;
                         a.nodes = t->d_nodes; a.tris = t->d_tris; a.trisf = t->d_trisf; a.mag = t->mag; a.perm = perm; a.counters = t->d_counters;
                         a.pts = (const double*)in; a.n = cnt; a.cp = (double*)o[0]; a.prim = (uint32_t*)o[1]; a.sqd = (double*)o[2]; a.flags = (uint8_t*)o[3];
                         a.cursor = t->warp[slot].cursor; a.spill = t->warp[slot].spill; a.slots64 = t->warp[slot].slots64; a.n_dev = nullptr;
                         a.exact_go = t->opt_warp_exact_go > 0 ? t->opt_warp_exact_go : 160;
                         a.flag_tie = B200AABB_FLAG_TIE; a.flag_exact = B200AABB_FLAG_EXACT;
                         const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)t->warp_grid, (cnt + 31) / 32);
                         if (t->segments) closest_point_warp_kernel<true><<<grid, 32, 0, st>>>(a);
                         else closest_point_warp_kernel<false><<<grid, 32, 0, st>>>(a);
                     });
}

int b200aabb_side_of_mesh(b200aabb_tree* t, const double* pts, uint64_t n, b200aabb_memspace space, int8_t* side, uint8_t* attempts) {
    return b200aabb_side_of_mesh_transformed(t, nullptr, pts, n, space, side, attempts);
}
int b200aabb_side_of_mesh_transformed(b200aabb_tree* t, const double* xf, const double* pts, uint64_t n, b200aabb_memspace space, int8_t* side, uint8_t* attempts) { B2_LOCK(t);
    if (!t || (n && (!pts || !side))) return B200AABB_ERR_INVALID_ARGUMENT;
    { const int rc_ = need_triangles(t, "side_of_mesh"); if (rc_) return rc_; }
    int rc = ensure_built(t);
    if (rc || n == 0) return rc;
    if (t->n == 0) {                                                   // no primitives: everything is outside
        if (space == B200AABB_HOST) { memset(side, 0xFF, n); if (attempts) memset(attempts, 0, n); return B200AABB_OK; }
        B2_CUDA(t, cudaSetDevice(t->device));
        B2_CUDA(t, cudaMemsetAsync(side, 0xFF, n, t->stream));
        if (attempts) B2_CUDA(t, cudaMemsetAsync(attempts, 0, n, t->stream));
        return B200AABB_OK;
    }
    B2_CUDA(t, cudaSetDevice(t->device));
    SideXf x;
    memset(&x, 0, sizeof(x));
    if (xf) {   // the image of the root box, conservatively (bvh.cuh::AffXf::box_meets does the same per node)
        memcpy(x.xf.m, xf, sizeof(x.xf.m));
        if (!image_of_box(xf, t->h_bbox, x.box)) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "side_of_mesh: the transformation is not finite");
        for (int k = 0; k < 6; ++k) x.mag = std::max(x.mag, fabs(x.box[k]));
    }
    return run_batch(t, pts, 24, n, space, kPointQuery, {{side, 1, 0}, {attempts, 1, 0}}, [&](cudaStream_t st, const void* in, void** o, uint64_t cnt, const uint32_t* perm) {
        if (xf) side_of_mesh_kernel<AffXf><<<grid_for(cnt, 128), 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, t->d_rootbox, (int8_t*)o[0], (uint8_t*)o[1], x);
        else side_of_mesh_kernel<NoXf><<<grid_for(cnt, 128), 128, 0, st>>>(tree_ref(t, perm), (const double*)in, cnt, t->d_rootbox, (int8_t*)o[0], (uint8_t*)o[1], x);
    });
}

int b200aabb_max_distance_to_mesh(b200aabb_tree* t, const double* pts, uint64_t n, b200aabb_memspace space, double* max_distance, uint64_t* farthest) {
    if (!t || (n && !pts) || !max_distance) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_LOCK(t);
    int rc = ensure_built(t);
    if (rc) return rc;
    if (t->n == 0) return fail(t, B200AABB_ERR_EMPTY_TREE, "distance query on an empty tree (CGAL_precondition(!empty()))");
    *max_distance = 0.0;
    if (farthest) *farthest = 0;
    if (n == 0) return B200AABB_OK;
    B2_CUDA(t, cudaSetDevice(t->device));
    double* d_pts = nullptr; double* d_sqd = nullptr; unsigned long long* d_best = nullptr;
    auto release = [&]() { if (space == B200AABB_HOST) cudaFree(d_pts); cudaFree(d_sqd); cudaFree(d_best); };
    cudaError_t e = cudaMalloc(&d_sqd, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_best, 16);
    if (e == cudaSuccess && space == B200AABB_HOST) e = cudaMalloc(&d_pts, n * 24);
    if (e != cudaSuccess) { release(); t->err = "cudaMalloc(max_distance_to_mesh)"; return B200AABB_ERR_OUT_OF_MEMORY; }
    if (space == B200AABB_HOST) cudaMemcpyAsync(d_pts, pts, n * 24, cudaMemcpyHostToDevice, t->stream);
    else d_pts = const_cast<double*>(pts);
    rc = b200aabb_closest_point(t, d_pts, n, B200AABB_DEVICE, nullptr, nullptr, d_sqd, nullptr);
    if (rc) { release(); return rc; }
    const unsigned long long init[2] = {0ull, ~0ull};
    cudaMemcpyAsync(d_best, init, 16, cudaMemcpyHostToDevice, t->stream);
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
    max_sqdist_kernel<<<grid, 256, 0, t->stream>>>(d_sqd, n, d_best);
    argmax_sqdist_kernel<<<grid, 256, 0, t->stream>>>(d_sqd, n, d_best, d_best + 1);
    t->stats.kernel_launches += 2;
    unsigned long long out[2] = {0, 0};
    cudaMemcpyAsync(out, d_best, 16, cudaMemcpyDeviceToHost, t->stream);
    e = cudaStreamSynchronize(t->stream);
    release();
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { t->err = "max_distance_to_mesh: CUDA failure"; return B200AABB_ERR_CUDA; }
    double sq; memcpy(&sq, &out[0], 8);
    *max_distance = sqrt(sq);                 // to_double(approximate_sqrt(sq_hdist)), distance.h:186
    if (farthest) *farthest = out[1] == ~0ull ? 0 : out[1];
    b200aabb_stats st;
    return b200aabb_get_stats(t, &st);        // (surfaces an incomplete traversal of the DEVICE batch)
}

int b200aabb_export_layout(b200aabb_tree* t, void** ptrs, uint64_t* bytes, int cap, int* count) { B2_LOCK(t);
    if (!t || !ptrs || !bytes || !count || cap < 4) return B200AABB_ERR_INVALID_ARGUMENT;
    { const int rc_ = need_triangles(t, "export_layout"); if (rc_) return rc_; }
    int rc = ensure_built(t);
    if (rc) return rc;
    ptrs[0] = t->d_nodes; bytes[0] = t->n > 1 ? (t->n - 1) * sizeof(BvhNode) : 0;
    ptrs[1] = t->d_tris; bytes[1] = t->n * sizeof(TriRec);
    ptrs[2] = t->d_rootbox; bytes[2] = 48;
    ptrs[3] = t->d_trisf; bytes[3] = t->n * sizeof(TriF);
    *count = 4;
    return B200AABB_OK;
}

int b200aabb_import_alloc(b200aabb_tree* t, uint64_t n) { B2_LOCK(t);
    if (!t) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = b200aabb_clear(t);
    if (rc) return rc;
    if (n == 0) return B200AABB_OK;
    B2_CUDA(t, cudaMalloc(&t->d_nodes, std::max<uint64_t>(n - 1, 1) * sizeof(BvhNode)));
    B2_CUDA(t, cudaMalloc(&t->d_tris, n * sizeof(TriRec)));
    B2_CUDA(t, cudaMalloc(&t->d_trisf, n * sizeof(TriF)));
    t->n = n; t->root = n == 1 ? (kLeafBit | 0u) : 0u; t->built = true; t->imported = true;
    return B200AABB_OK;
}

int b200aabb_import_finish(b200aabb_tree* t) { B2_LOCK(t);
    if (!t) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_CUDA(t, cudaSetDevice(t->device));
    if (t->n) B2_CUDA(t, cudaMemcpy(t->h_bbox, t->d_rootbox, 48, cudaMemcpyDeviceToHost));
    t->mag = 0.0;
    for (int a = 0; a < 6; ++a) if (t->n) t->mag = std::max(t->mag, fabs(t->h_bbox[a]));
    return B200AABB_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Multi-GPU behind the C ABI: the tree is built on one rank and replicated with NCCL over NVLink / NVSwitch; every rank
// then answers its own block of the query batch (no collective in the query loop).  NCCL is resolved at run time (dlopen), so the
// library loads — and every single-GPU entry point works — without it.
// ------------------------------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // the soname first: a process that already carries an NCCL (e.g. PyTorch's) gets that very copy
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) return;
        auto sym = [&](const char* n) { return dlsym(api.lib, n); };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.Broadcast && api.GroupStart && api.GroupEnd;
    });
    return api;
}

// binary32 copies of the primitive records (the receiver of a broadcast derives them instead of receiving 64 bytes per primitive)
__global__ void __launch_bounds__(256) derive_trisf_kernel(const TriRec* __restrict__ tris, uint64_t n, TriF* __restrict__ outf) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    TriF f;
#pragma unroll
    for (int j = 0; j < 9; ++j) f.v[j] = (float)tris[k].v[j];
#pragma unroll
    for (int j = 0; j < 7; ++j) f.pad[j] = 0;
    outf[k] = f;
}

}  // namespace

struct b200aabb_comm {
    ncclComm_t comm = nullptr;
    int device = 0, rank = 0, world = 1;
    unsigned long long* d_header = nullptr;   // [4]: primitives, has ids, mode, reserved
    std::string err;
};

namespace {

#define B2_NCCL(c, call)                                                                                        \
    do {                                                                                                         \
        ncclResult_t r_ = (call);                                                                                \
        if (r_ != ncclSuccess) {                                                                                 \
            (c)->err = std::string(#call) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "NCCL error"); \
            return B200AABB_ERR_NCCL;                                                                            \
        }                                                                                                        \
    } while (0)

// the per-rank part of a tree replication; the caller brackets it with ncclGroupStart / End when it drives several ranks from one thread
// (phase 0: header, phase 1: payload — the payload sizes depend on the header, so the two can not share a group)
int bcast_header(b200aabb_comm* c, b200aabb_tree* t, int root, int mode, cudaStream_t st) {
    if (c->rank == root) {
        const unsigned long long h[4] = {t->n, t->d_ids ? 1ull : 0ull, (unsigned long long)mode, t->segments ? 1ull : 0ull};
        if (cudaMemcpyAsync(c->d_header, h, sizeof(h), cudaMemcpyHostToDevice, st) != cudaSuccess) return B200AABB_ERR_CUDA;
        if (cudaStreamSynchronize(st) != cudaSuccess) return B200AABB_ERR_CUDA;   // (h lives on this stack frame)
    }
    B2_NCCL(c, nccl().Broadcast(c->d_header, c->d_header, 32, ncclUint8, root, c->comm, st));
    return B200AABB_OK;
}

int bcast_prepare(b200aabb_comm* c, b200aabb_tree* t, int root, int mode, cudaStream_t st) {
    unsigned long long h[4];
    if (cudaMemcpyAsync(h, c->d_header, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return B200AABB_ERR_CUDA;
    if (c->rank == root) return B200AABB_OK;
    const uint64_t n = h[0];
    const bool segments = h[3] != 0;                 // the kind of primitive travels with the header (the receiving calls below reset it)
    if (mode == 0) {
        if (t->imported && t->n == n && t->d_tris && t->d_trisf && (n < 2 || t->d_nodes)) { t->segments = segments; return B200AABB_OK; }   // a replica of the same size: refresh it in place
        const int rc = b200aabb_import_alloc(t, n);
        t->segments = segments;
        return rc;
    }
    // (a handle that already holds inserted primitives of the same size — an earlier mode-1 replica — receives into its buffers: a re-build then allocates nothing)
    if (!t->imported && t->n == n && n && t->d_soup && (h[1] != 0) == (t->d_ids != nullptr)) { t->built = false; t->bounds_valid = false; t->segments = segments; return B200AABB_OK; }
    int rc = b200aabb_clear(t);
    t->segments = segments;
    if (rc || n == 0) return rc;
    if (cudaMalloc(&t->d_soup, n * 72 + kSoupSlack) != cudaSuccess) return B200AABB_ERR_OUT_OF_MEMORY;
    if (h[1] && cudaMalloc(&t->d_ids, n * 4) != cudaSuccess) return B200AABB_ERR_OUT_OF_MEMORY;
    t->n = n; t->built = false; t->bounds_valid = false;
    return B200AABB_OK;
}

int bcast_payload(b200aabb_comm* c, b200aabb_tree* t, int root, int mode, cudaStream_t st, uint64_t* bytes) {
    const uint64_t n = t->n;
    *bytes = 0;
    if (n == 0) return B200AABB_OK;
    if (mode == 0) {
        if (n > 1) { B2_NCCL(c, nccl().Broadcast(t->d_nodes, t->d_nodes, (n - 1) * sizeof(BvhNode), ncclUint8, root, c->comm, st)); *bytes += (n - 1) * sizeof(BvhNode); }
        B2_NCCL(c, nccl().Broadcast(t->d_tris, t->d_tris, n * sizeof(TriRec), ncclUint8, root, c->comm, st));
        B2_NCCL(c, nccl().Broadcast(t->d_rootbox, t->d_rootbox, 48, ncclUint8, root, c->comm, st));
        *bytes += n * sizeof(TriRec) + 48;
    } else {
        B2_NCCL(c, nccl().Broadcast(t->d_soup, t->d_soup, n * 72, ncclUint8, root, c->comm, st));
        *bytes += n * 72;
        if (t->d_ids) { B2_NCCL(c, nccl().Broadcast(t->d_ids, t->d_ids, n * 4, ncclUint8, root, c->comm, st)); *bytes += n * 4; }
    }
    return B200AABB_OK;
}

int bcast_finish(b200aabb_comm* c, b200aabb_tree* t, int root, int mode, cudaStream_t st) {
    if (c->rank == root || t->n == 0) return B200AABB_OK;
    if (mode == 0) {
        derive_trisf_kernel<<<grid_for(t->n, 256), 256, 0, st>>>(t->d_tris, t->n, t->d_trisf);
        t->stats.kernel_launches += 1;
        if (cudaGetLastError() != cudaSuccess) return B200AABB_ERR_CUDA;
        return B200AABB_OK;   // (b200aabb_import_finish after the stream has drained: it reads the root box)
    }
    return do_build(t);
}

}  // namespace

extern "C" {

int b200aabb_comm_unique_id(char id[128]) {
    if (!id) return B200AABB_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return B200AABB_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != ncclSuccess) return B200AABB_ERR_NCCL;
    memcpy(id, &u, 128);
    return B200AABB_OK;
}

static int comm_setup(b200aabb_comm* c) {
    if (cudaSetDevice(c->device) != cudaSuccess) return B200AABB_ERR_NO_DEVICE;
    if (cudaMalloc(&c->d_header, 32) != cudaSuccess) return B200AABB_ERR_OUT_OF_MEMORY;
    cudaMemset(c->d_header, 0, 32);
    return B200AABB_OK;
}

int b200aabb_comm_init(int device, int rank, int world, const char id[128], b200aabb_comm** out) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return B200AABB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!nccl().ok) return B200AABB_ERR_NCCL;
    b200aabb_comm* c = new b200aabb_comm();
    c->device = device; c->rank = rank; c->world = world;
    int rc = comm_setup(c);
    if (rc) { delete c; return rc; }
    ncclUniqueId u;
    memcpy(&u, id, 128);
    if (nccl().CommInitRank(&c->comm, world, u, rank) != ncclSuccess) { cudaFree(c->d_header); delete c; return B200AABB_ERR_NCCL; }
    // warm-up: the first collective of a communicator sets up its channels (tens of milliseconds); do it here, not in the first tree replication
    if (nccl().Broadcast(c->d_header, c->d_header, 32, ncclUint8, 0, c->comm, nullptr) != ncclSuccess || cudaStreamSynchronize(nullptr) != cudaSuccess) {
        b200aabb_comm_destroy(c);
        return B200AABB_ERR_NCCL;
    }
    *out = c;
    return B200AABB_OK;
}

int b200aabb_comm_init_all(int ndev, const int* devices, b200aabb_comm** out) {
    if (!out || ndev < 1 || ndev > 64) return B200AABB_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < ndev; ++i) out[i] = nullptr;
    if (!nccl().ok) return B200AABB_ERR_NCCL;
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
    std::vector<ncclComm_t> comms(ndev);
    if (nccl().CommInitAll(comms.data(), ndev, devs.data()) != ncclSuccess) return B200AABB_ERR_NCCL;
    int rc = B200AABB_OK;
    for (int i = 0; i < ndev; ++i) {
        out[i] = new b200aabb_comm();
        out[i]->comm = comms[i]; out[i]->device = devs[i]; out[i]->rank = i; out[i]->world = ndev;
        if (int r = comm_setup(out[i])) rc = r;
    }
    if (rc == B200AABB_OK) {   // warm-up, see b200aabb_comm_init
        nccl().GroupStart();
        for (int i = 0; i < ndev; ++i) { cudaSetDevice(devs[i]); nccl().Broadcast(out[i]->d_header, out[i]->d_header, 32, ncclUint8, 0, comms[i], nullptr); }
        if (nccl().GroupEnd() != ncclSuccess) rc = B200AABB_ERR_NCCL;
        for (int i = 0; i < ndev; ++i) { cudaSetDevice(devs[i]); if (cudaStreamSynchronize(nullptr) != cudaSuccess) rc = B200AABB_ERR_CUDA; }
    }
    if (rc) for (int i = 0; i < ndev; ++i) { b200aabb_comm_destroy(out[i]); out[i] = nullptr; }
    return rc;
}

int b200aabb_comm_destroy(b200aabb_comm* c) {
    if (!c) return B200AABB_OK;
    cudaSetDevice(c->device);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    cudaFree(c->d_header);
    delete c;
    return B200AABB_OK;
}

const char* b200aabb_comm_last_error(const b200aabb_comm* c) { return c ? c->err.c_str() : "null communicator"; }

int b200aabb_comm_broadcast_tree(b200aabb_comm* c, b200aabb_tree* t, int root, int mode, double* ms, uint64_t* bytes) {
    if (!c || !t || root < 0 || root >= c->world || (mode != 0 && mode != 1) || t->device != c->device) return B200AABB_ERR_INVALID_ARGUMENT;
    B2_LOCK(t);
    B2_CUDA(t, cudaSetDevice(t->device));
    cudaStream_t st = t->stream;
    int rc = B200AABB_OK;
    if (c->rank == root) {
        rc = ensure_built(t);
        if (rc) return rc;
        if (mode == 1 && t->n && !t->d_soup) return fail(t, B200AABB_ERR_INVALID_ARGUMENT, "mode 1 replicates the inserted primitives: the root holds none (imported tree)");
    }
    // (events of its own: the build that mode 1 runs on the receivers records the handle's ev_a / ev_b itself)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    B2_CUDA(t, cudaEventCreate(&e0));
    B2_CUDA(t, cudaEventCreate(&e1));
    auto done = [&](int code) { cudaEventDestroy(e0); cudaEventDestroy(e1); return code; };
    if (cudaEventRecord(e0, st) != cudaSuccess) return done(B200AABB_ERR_CUDA);
    if ((rc = bcast_header(c, t, root, mode, st))) return done(rc);
    if ((rc = bcast_prepare(c, t, root, mode, st))) return done(rc);
    uint64_t nbytes = 0;
    if ((rc = bcast_payload(c, t, root, mode, st, &nbytes))) return done(rc);
    if ((rc = bcast_finish(c, t, root, mode, st))) return done(rc);
    if (cudaEventRecord(e1, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return done(B200AABB_ERR_CUDA);
    if (c->rank != root && mode == 0) { rc = b200aabb_import_finish(t); if (rc) return done(rc); }
    float f = 0; cudaEventElapsedTime(&f, e0, e1);
    done(0);
    if (ms) *ms = f;
    if (bytes) *bytes = nbytes;
    return B200AABB_OK;
}

int b200aabb_comm_broadcast_tree_all(b200aabb_comm** cs, b200aabb_tree** ts, int ndev, int root, int mode, double* ms, uint64_t* bytes) {
    if (!cs || !ts || ndev < 1 || root < 0 || root >= ndev || (mode != 0 && mode != 1)) return B200AABB_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < ndev; ++i) if (!cs[i] || !ts[i] || ts[i]->device != cs[i]->device || cs[i]->world != ndev) return B200AABB_ERR_INVALID_ARGUMENT;
    int rc = ensure_built(ts[root]);
    if (rc) return rc;
    if (mode == 1 && ts[root]->n && !ts[root]->d_soup) return fail(ts[root], B200AABB_ERR_INVALID_ARGUMENT, "mode 1 replicates the inserted primitives: the root holds none");
    std::vector<cudaEvent_t> e0(ndev, nullptr), e1(ndev, nullptr);     // (events of their own: mode 1's build records the handles' ev_a / ev_b)
    struct Cleanup { std::vector<cudaEvent_t>&a, &b; b200aabb_tree** ts; ~Cleanup() { for (size_t i = 0; i < a.size(); ++i) { cudaSetDevice(ts[i]->device); if (a[i]) cudaEventDestroy(a[i]); if (b[i]) cudaEventDestroy(b[i]); } } } cleanup{e0, e1, ts};
    for (int i = 0; i < ndev; ++i) { cudaSetDevice(ts[i]->device); if (cudaEventCreate(&e0[i]) != cudaSuccess || cudaEventCreate(&e1[i]) != cudaSuccess) return B200AABB_ERR_CUDA; cudaEventRecord(e0[i], ts[i]->stream); }
    nccl().GroupStart();
    for (int i = 0; i < ndev && !rc; ++i) { cudaSetDevice(ts[i]->device); rc = bcast_header(cs[i], ts[i], root, mode, ts[i]->stream); }
    if (nccl().GroupEnd() != ncclSuccess && !rc) rc = B200AABB_ERR_NCCL;
    for (int i = 0; i < ndev && !rc; ++i) { cudaSetDevice(ts[i]->device); rc = bcast_prepare(cs[i], ts[i], root, mode, ts[i]->stream); }
    if (rc) return rc;
    uint64_t nbytes = 0;
    nccl().GroupStart();
    for (int i = 0; i < ndev && !rc; ++i) { cudaSetDevice(ts[i]->device); rc = bcast_payload(cs[i], ts[i], root, mode, ts[i]->stream, &nbytes); }
    if (nccl().GroupEnd() != ncclSuccess && !rc) rc = B200AABB_ERR_NCCL;
    for (int i = 0; i < ndev && !rc; ++i) { cudaSetDevice(ts[i]->device); rc = bcast_finish(cs[i], ts[i], root, mode, ts[i]->stream); }
    if (rc) return rc;
    double worst = 0;
    for (int i = 0; i < ndev; ++i) {
        cudaSetDevice(ts[i]->device);
        cudaEventRecord(e1[i], ts[i]->stream);
        if (cudaStreamSynchronize(ts[i]->stream) != cudaSuccess) return B200AABB_ERR_CUDA;
        if (i != root && mode == 0) { rc = b200aabb_import_finish(ts[i]); if (rc) return rc; }
        float f = 0; cudaEventElapsedTime(&f, e0[i], e1[i]);
        worst = std::max(worst, (double)f);
    }
    if (ms) *ms = worst;
    if (bytes) *bytes = nbytes;
    return B200AABB_OK;
}

}  // extern "C"
