// closest_warp.cuh — warp-cooperative closest_point_and_primitive(Point_3): one persistent warp serves a BATCH of 32 queries
// (one per "slot") and keeps ONE shared work pool for all of them in shared memory.
//
// Replaces the same reference path as bvh.cuh::traverse_closest (AABB_tree.h:987-1055, Projection_traits
// AABB_traversal_traits.h:277-327, AABB_traits_3.h:715-733) and produces the same results: the winner is the candidate
// (reference-rounded projection, geom.cuh::closest_on_triangle) with the smallest EXACT squared distance, ties -> smallest
// primitive id + FLAG_TIE.  Only the scheduling differs.
//
// Why (profiles/r1_03): in the thread-per-query kernel a closest_point query visits ~170 nodes and ~75 leaves, the counts
// differ a lot between the 32 queries of a warp, and each thread runs its own three-phase loop: 5.2 of 32 threads are active
// per issued instruction.  Here the unit of work is a (slot, node) or (slot, leaf) PAIR, the pairs of the 32 slots are pooled,
// and every phase takes 32 pairs off a pool whichever slots they belong to:
//   (A)  node phase     pop 32 (slot,node) entries (LIFO: depth-first per slot); drop entries whose box has fallen outside the
//                       slot's ball; load the node, test both child boxes against the slot's binary32 point (directed rounding);
//                       push internal children (far first, near on top) and record leaf children as (slot,leaf) pairs
//   (B1) pre-test 1     rigorous binary32 LOWER bound of the distance of 32 (slot,leaf) pairs (along the centroid direction) on a
//                       64-byte binary32 copy of the triangle; pairs certainly outside the ball retire
//   (B2) pre-test 2     approximate closest point of the triangle in binary32 (divisions, branches): a sharper lower bound, and a
//                       rigorous UPPER bound of the distance that shrinks the slot's ball at once (atomicMin), before any binary64 work
//   (B3) exact phase    pairs still inside the ball: the reference's Construct_projected_point_3 in binary64, then the exact
//                       Compare_distance_3 against the slot's incumbent (all candidates in parallel, one replacement per slot and round)
// Each stage has its own pool and runs when 32 pairs wait (or when nothing else is left), so every stage runs with (nearly) all
// lanes busy; a pair's stored lower bound is re-checked against the slot's current ball at every hand-over for free.
// A slot's state (point, incumbent, bounds) lives in shared memory, packed so that a stage reads it with one or two 128-bit loads.
// When all pools are empty the batch is complete: the lanes write their slots' results and the warp takes the next 32 queries
// from the batch cursor (queries are handed out in coherence order).  A new query first walks down to its nearest leaf on its
// own lane (nearest child first), pushing the far children, and projects on that leaf: that is its first ball (the kd-tree hint
// of the reference, AABB_tree.h:591-612, is not needed).
// The node pool is a stack in shared memory whose bottom half is spilled to a per-warp area in global memory when it fills
// and reloaded when the stack runs empty, so depth is bounded by kWpSpillCap only (overflow is reported, never silent).
// What bounds this kernel (profiles/r1_05 .. r1_15): the L1 data stage, one wavefront per cycle, because every lane fetches its own node /
// triangle: hence the packed slot state, the binary32 triangle copy and no per-slot bookkeeping atomics.
#pragma once
#include "bvh.cuh"

namespace b2 {

#ifndef B2_WP_NODE_CAP
#define B2_WP_NODE_CAP 768
#endif
#ifdef B2_WP_NO_STATS               // experiment (profiles/r2_29): the per-round traversal counters compiled out
#define B2_WP_STAT(x)
#else
#define B2_WP_STAT(x) x
#endif
constexpr int kWpNodeCap = B2_WP_NODE_CAP;                 // (slot,node) entries held in shared memory
constexpr int kWpSpillChunk = kWpNodeCap / 2;   // entries moved per spill / reload
constexpr int kWpSpillCap = 16 * kWpSpillChunk; // entries per warp in the global spill area
constexpr int kWpLeafCap = 96;                  // < 32 waiting + at most 64 added by one node phase
constexpr int kWpMidCap = 128;                  // < 32 waiting + at most 95 added by one pre-test-1 phase
constexpr int kWpExactCap = 192;                // pre-test 2 stops handing over when fewer than 32 places are left
#ifndef B2_WP_WARPS_PER_SM
#define B2_WP_WARPS_PER_SM 20
#endif
constexpr int kWpWarpsPerSM = B2_WP_WARPS_PER_SM;               // resident warps per SM this kernel is sized for (registers <= 102, shared memory <= 10.4 KB): measured best of 16 / 20 / 24 / 28
constexpr uint32_t kWpRefMask = 0x07FFFFFFu;    // 27 bits of leaf slot next to the 5-bit query slot
constexpr uint64_t kWpMaxPrims = 1ull << 27;

struct WarpShared {
    // (npool / lpool end in a SINK entry: lanes with nothing to push store there instead of branching; the sink is never read, so the
    //  write/write hazards compute-sanitizer's racecheck reports on it are the only ones it reports for this kernel, and are benign)
    uint2 npool[kWpNodeCap + 1];  // x = child reference (node index, or kLeafBit | leaf slot); y = lower bound of the box's squared distance from
                                  //     the binary32 query point (low 5 mantissa bits cleared = rounded down) | query slot.  [+1: sink of masked pushes]
    uint32_t lpool[kWpLeafCap + 1];   // query slot << 27 | leaf slot : awaiting pre-test 1
    uint2 mpool[kWpMidCap];       // x = same, y = lower bound of the distance (binary32) : awaiting pre-test 2
    uint2 epool[kWpExactCap];     // same : awaiting the binary64 projection
    float4 qa[32];                // query point rounded to binary32 (x,y,z); w = (rbound + perr)^2 rounded up (node phase)
    float4 qc[32];                // x = rbound: pruning radius >= sqrt(widened squared distance of the eventual winner);
                                  // y = e2 >= |error of a binary32 relative vertex| = sqrt(3) ev + perr; z = dslack >= 2 delta; w = perr >= |query - qa.xyz|
    uint32_t bprim[32];           // incumbent primitive id
    uint32_t qidx[32];            // the query's index in the caller's order
    uint32_t flags[32];           // B200AABB_FLAG_TIE / _EXACT
};

// binary64 state of the 32 slots: only the exact phase, the refill and the result write touch it, so it lives in global memory (L2)
// and leaves the shared memory to the pools (more resident warps)
struct WarpSlots64 {
    double o[3][32];              // query point
    double bp[3][32];             // incumbent closest point
    double bsqd[32];              // its squared distance as the reference computes it
};

struct WarpClosestArgs {
    const BvhNode* nodes;
    const TriRec* tris;
    const TriF* trisf;             // binary32 copy of the triangles (same order), for the pre-tests
    double mag;                    // largest coordinate magnitude of the root box
    const uint32_t* perm;          // coherence order (slot i of the order serves query perm[i]); nullptr = identity
    unsigned long long* counters;  // as TreeRef::counters
    const double* pts;
    uint64_t n;
    double* cp;
    uint32_t* prim;
    double* sqd;
    uint8_t* flags;
    unsigned long long* cursor;    // next position of the batch to hand out (zeroed before the launch)
    uint2* spill;                  // gridDim.x * kWpSpillCap entries
    WarpSlots64* slots64;          // gridDim.x records
    const unsigned long long* n_dev; // when not null: the number of queries is read from here (the overflow list of closest_filter_kernel)
    int exact_go;                  // run the exact phase when at least this many pairs wait (1 .. kWpExactCap - 32)
    uint32_t flag_tie, flag_exact;
};

#if defined(__CUDACC__)

// LOWER bound of the squared distance from the binary32 point q to the closed box (every operation rounds towards the smaller
// result, as bvh.cuh::point_box_sqdist).  The query itself is within perr of q, so a box can only matter when this value is
// <= (rbound + perr)^2 = qa.w.
__device__ __forceinline__ float wp_box_sqdist(const float* lo, const float* hi, const float4& q) {
    const float dx = fmaxf(fmaxf(__fsub_rd(lo[0], q.x), __fsub_rd(q.x, hi[0])), 0.0f);
    const float dy = fmaxf(fmaxf(__fsub_rd(lo[1], q.y), __fsub_rd(q.y, hi[1])), 0.0f);
    const float dz = fmaxf(fmaxf(__fsub_rd(lo[2], q.z), __fsub_rd(q.z, hi[2])), 0.0f);
    return __fmaf_rd(dz, dz, __fmaf_rd(dy, dy, __fmul_rd(dx, dx)));
}

// UPPER bound of the squared distance from q to the farthest corner of the box (every operation rounds up)
__device__ __forceinline__ float wp_box_far_sqdist(const float* lo, const float* hi, const float4& q) {
    const float fx = fmaxf(__fsub_ru(q.x, lo[0]), __fsub_ru(hi[0], q.x));
    const float fy = fmaxf(__fsub_ru(q.y, lo[1]), __fsub_ru(hi[1], q.y));
    const float fz = fmaxf(__fsub_ru(q.z, lo[2]), __fsub_ru(hi[2], q.z));
    return __fmaf_ru(fz, fz, __fmaf_ru(fy, fy, __fmul_ru(fx, fx)));
}

// shrinks slot `owner`'s ball to radius rb (any lane, any time: both bounds only ever decrease)
__device__ __forceinline__ void wp_shrink(WarpShared& S, unsigned owner, float rb) {
    atomicMin(reinterpret_cast<unsigned*>(&S.qc[owner].x), __float_as_uint(rb));
    const float r = __fadd_ru(rb, S.qc[owner].w);
    atomicMin(reinterpret_cast<unsigned*>(&S.qa[owner].w), __float_as_uint(__fmul_ru(r, r)));
}

// pruning radius of a slot from its incumbent's squared distance: rbound >= sqrt(widen_sqdist(sqd, delta)) (bvh.cuh), with the
// square roots taken in binary32 rounded up (s >= sqrt(sqd) makes the widened value only larger).  Unless `reset`, the ball only
// shrinks: the binary32 upper bounds of wp_sharp may already be tighter than this incumbent.
__device__ __forceinline__ void wp_set_bounds(WarpShared& S, unsigned owner, double sqd, const Vec3& /*v*/, double delta, bool reset) {
    const float s = __fsqrt_ru(float_above(sqd));
    const float rb = __fsqrt_ru(float_above(sqd + ((2.0 * delta * (double)s + delta * delta) + sqd * 1e-13)));
    if (reset) {
        S.qc[owner].x = rb;
        const float r = __fadd_ru(rb, S.qc[owner].w);
        S.qa[owner].w = __fmul_ru(r, r);
    } else wp_shrink(S, owner, rb);
}

// Binary32 bounds of the distance from the query to one primitive, on coordinates RELATIVE to the query computed from the binary32
// triangle copy and the binary32 query point.  Each relative coordinate is off by <= m 2^-24 (the subtraction; m = largest
// magnitude among them) + the rounding of the vertex (<= ev = 2^-24 * largest scene coordinate) + that of the query: as vectors,
// every relative vertex is within e2 = sqrt(3) ev + perr of its true position (plus the m 2^-24 part, which wp_lower / wp_sharp
// carry themselves), so a distance bound computed from them moves by at most e2.
struct RelTri { float ra[3], rb[3], rc[3], m; };
__device__ __forceinline__ RelTri wp_relative(const TriF* __restrict__ trisf, uint32_t slot, const float4& q) {
    uint32_t f[8];
    ldg256(trisf + slot, f);
    const float cz = __ldg(&trisf[slot].v[8]);
    RelTri r;
    r.ra[0] = __uint_as_float(f[0]) - q.x; r.ra[1] = __uint_as_float(f[1]) - q.y; r.ra[2] = __uint_as_float(f[2]) - q.z;
    r.rb[0] = __uint_as_float(f[3]) - q.x; r.rb[1] = __uint_as_float(f[4]) - q.y; r.rb[2] = __uint_as_float(f[5]) - q.z;
    r.rc[0] = __uint_as_float(f[6]) - q.x; r.rc[1] = __uint_as_float(f[7]) - q.y; r.rc[2] = cz - q.z;
    r.m = fmaxf(fmaxf(fmaxf(fabsf(r.ra[0]), fabsf(r.ra[1])), fmaxf(fabsf(r.ra[2]), fabsf(r.rb[0]))),
                fmaxf(fmaxf(fabsf(r.rb[1]), fabsf(r.rb[2])), fmaxf(fmaxf(fabsf(r.rc[0]), fabsf(r.rc[1])), fabsf(r.rc[2]))));
    return r;
}
// LOWER bound from a direction u (bvh.cuh::certainly_farther solved for the distance): every point x of the triangle has
// |x| >= min_i (u.r_i)/|u|; E = 5e-7 m |u| covers the coordinate rounding and the dot products, the factor 0.999998 the rounding of
// |u|, of its reciprocal and of the two products.  -inf when u is (numerically) zero.  (Relative to the binary32 data: the caller subtracts e2.)
B2_HD float wp_lower(const float* u, float su, const RelTri& r) {   // (host build: tests/hostcheck checks the bounds against exact distances)
    const float md = fminf(dot3f(u, r.ra), fminf(dot3f(u, r.rb), dot3f(u, r.rc)));
#if defined(__CUDA_ARCH__)
    const float rcp = __frcp_rn(su);
#else
    const float rcp = 1.0f / su;
#endif
    // Range guard (binary32 under/overflow): the bound is only formed when m < 1e18 (no product can overflow: |u.r_i| <= 3 m |u| < 1e37) and
    // m |u| > 1e-30 (the dot products and the allowance E are far above the denormal range, whose absolute rounding would exceed E).
    // Outside that range: no bound (the pair goes on to the binary64 stage), never a wrong one.
    const bool in_range = su > 1e-18f && su < 1e19f && r.m < 1e18f && r.m * su > 1e-30f;   // (su = inf / NaN: the direction itself is the product of an overflow)
    return in_range ? fmaf(-5e-7f * r.m, su, md) * (rcp * 0.999998f) : -INFINITY;
}
// Approximate closest point P = a + v ab + w ac (v, w >= 0, v + w <= 1 up to an ulp) -> a sharper lower bound (direction P) and a
// rigorous UPPER bound of (distance + the slot's rounding allowance): P's coordinates are within 1.2e-6 m of a point of the true
// triangle (vertex rounding m 2^-24 each, six binary32 operations on magnitudes <= 2 m), so dist <= |P| + 2.1e-6 m; the formula
// allows 8e-6 m + 1e-5 |P| + dslack.  The winner's candidate is no farther than this primitive's candidate, which lies within the
// rounding allowance of the true distance (the premise of widen_sqdist), so pruning with `hi` never discards the winner or a
// primitive tied with it.  hi = NaN (no bound) compares false everywhere.  e2: see wp_relative.
B2_HD void wp_sharp(const RelTri& r, float dslack, float e2, float* lower, float* hi) {
    float u[3];
    approx_closest_to_origin(r.ra, r.rb, r.rc, u);
    const float su = sqrtf(dot3f(u, u));
    *lower = wp_lower(u, su, r) - e2;
    // The upper bound may only be used when the reference's own projection on this primitive is accurate, i.e. when the triangle is
    // well conditioned: |ab x ac|^2 >= 1e-6 |ab|^2 |ac|^2 (sine of the angle at a >= 1e-3) with both edges well above the noise of
    // the binary32 coordinates.  For degenerate triangles and slivers Construct_projected_point_3 can return a point far from the
    // closest one (function_objects.h:3208-3247), so their candidates promise nothing: no bound.
    const float ab[3] = {r.rb[0] - r.ra[0], r.rb[1] - r.ra[1], r.rb[2] - r.ra[2]}, ac[3] = {r.rc[0] - r.ra[0], r.rc[1] - r.ra[1], r.rc[2] - r.ra[2]};
    const float n[3] = {ab[1] * ac[2] - ab[2] * ac[1], ab[2] * ac[0] - ab[0] * ac[2], ab[0] * ac[1] - ab[1] * ac[0]};
    const float lab = dot3f(ab, ab), lac = dot3f(ac, ac), noise = 1e3f * (e2 + r.m * 2.4e-7f);
    // (m in [1e-9, 1e9]: |n|^2 and |ab|^2 |ac|^2 ~ m^4 stay normal binary32 numbers, so the conditioning test and |u| itself mean what they say)
    const bool sound = r.m > 1e-9f && r.m < 1e9f && dot3f(n, n) >= 1e-6f * lab * lac && lab > noise * noise && lac > noise * noise;
#if defined(__CUDA_ARCH__)
    const float no_bound = __int_as_float(0x7fc00000);
#else
    const float no_bound = NAN;
#endif
    *hi = sound ? ((su * 1.00001f + 8e-6f * r.m) + dslack) + e2 : no_bound;
}

// The binary64 projection, out of line and with its arguments in registers: the kernel calls it from the refill and from the exact
// phase, and keeping this long routine (and its own register needs) out of the body keeps the node loop's allocation tight.
struct CandOut { double x, y, z; int nexact; };
// seg: the primitives are Segment_3 kept as (source, target, target) records (b200aabb_set_segments): the candidate is the reference's projection on the SEGMENT;
// the binary32 stages need no change (their bounds hold for the flat triangle = the segment; it yields no upper bound, so only exact candidates shrink the ball)
// (a template parameter, not an argument: one more live value across the call cost the triangle kernel 4.5 % — 88 -> 112 bytes of spills, measured)
template <bool SEG = false>
__device__ __noinline__ CandOut wp_candidate(double ax, double ay, double az, double bx, double by, double bz, double cx, double cy, double cz,
                                             double ox, double oy, double oz) {
    int nexact = 0;
    const Vec3 p = SEG ? closest_candidate_segment(Vec3{ax, ay, az}, Vec3{bx, by, bz}, Vec3{ox, oy, oz}, nexact)
                       : closest_candidate(Tri3{{ax, ay, az}, {bx, by, bz}, {cx, cy, cz}}, Vec3{ox, oy, oz}, nexact);
    return CandOut{p.x, p.y, p.z, nexact};
}

// the replacement step of the reference (AABB_traits_3.h:724-732) for slot `owner`, given sg = Compare_distance_3(query, candidate,
// incumbent) <= 0; same tie rule as bvh.cuh::traverse_closest
__device__ __forceinline__ void wp_apply(WarpShared& S, WarpSlots64& G, unsigned owner, int sg, const Vec3& cpt, const Vec3& v, double sqd, double delta, uint32_t id,
                                         uint32_t flag_tie) {
    uint32_t fl = S.flags[owner];
    bool take;
    if (sg == 0) { take = id < S.bprim[owner]; fl |= flag_tie; }
    else { take = true; fl &= ~flag_tie; }
    S.flags[owner] = fl;
    if (take) {
        __stcg(&G.bp[0][owner], cpt.x); __stcg(&G.bp[1][owner], cpt.y); __stcg(&G.bp[2][owner], cpt.z);
        __stcg(&G.bsqd[owner], sqd);
        S.bprim[owner] = id;
        wp_set_bounds(S, owner, sqd, v, delta, false);
    }
}

template <bool SEG>
__global__ void __launch_bounds__(32, kWpWarpsPerSM) closest_point_warp_kernel(const WarpClosestArgs a) {
    __shared__ WarpShared S;
    const unsigned lane = threadIdx.x;
    const unsigned full = 0xFFFFFFFFu;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned long long total = a.n_dev ? *a.n_dev : a.n;
    uint2* const myspill = a.spill + (size_t)blockIdx.x * kWpSpillCap;
    WarpSlots64& G = a.slots64[blockIdx.x];
    const float ev = float_above(a.mag * 5.9604645e-8 * 1.000001);   // >= rounding of a vertex coordinate to binary32
    int ncnt = 0, lcnt = 0, mcnt = 0, ecnt = 0, spilled = 0;         // warp-uniform
    unsigned active = 0;                                              // slots holding a query of the current batch
    // statistics as warp-uniform counts taken from ballots (per-lane counters cost five registers that the exact phase needs)
    uint32_t w_nodes = 0, w_leaves = 0, w_evals = 0, w_exact = 0, w_ties = 0;

    // moves the bottom kWpSpillChunk entries of the stack to global memory; false = spill area exhausted
    // moves the bottom kWpSpillChunk entries of the stack to global memory and the rest down; false = spill area exhausted.
    // (A ring buffer that avoids the move was measured 2 % slower: its index arithmetic sits in the node loop, the move is rare.)
    auto spill_bottom = [&]() -> bool {
        if (spilled + kWpSpillChunk > kWpSpillCap) return false;
        for (int i = lane; i < kWpSpillChunk; i += 32) myspill[spilled + i] = S.npool[i];
        __syncwarp();
        for (int base = kWpSpillChunk; base < ncnt; base += 32) {
            const int idx = base + (int)lane;
            uint2 v = make_uint2(0, 0);
            if (idx < ncnt) v = S.npool[idx];
            __syncwarp();
            if (idx < ncnt) S.npool[idx - kWpSpillChunk] = v;
            __syncwarp();
        }
        spilled += kWpSpillChunk; ncnt -= kWpSpillChunk;
        return true;
    };

    while (true) {
        if (ncnt == 0 && spilled > 0) {                       // reload the most recently spilled chunk
            for (int i = lane; i < kWpSpillChunk; i += 32) S.npool[i] = __ldcg(&myspill[spilled - kWpSpillChunk + i]);
            spilled -= kWpSpillChunk; ncnt = kWpSpillChunk;
            __syncwarp();
        }
        if (ncnt == 0 && lcnt == 0 && mcnt == 0 && ecnt == 0) {
            // -------------------------------------------------------------------------------------------- batch complete: results out, next 32 queries in
            if ((active >> lane) & 1u) {
                const uint64_t qi = S.qidx[lane];
                const uint32_t fl = S.flags[lane];
                if (a.cp) { a.cp[3 * qi] = __ldcg(&G.bp[0][lane]); a.cp[3 * qi + 1] = __ldcg(&G.bp[1][lane]); a.cp[3 * qi + 2] = __ldcg(&G.bp[2][lane]); }
                if (a.prim) a.prim[qi] = S.bprim[lane];
                if (a.sqd) a.sqd[qi] = __ldcg(&G.bsqd[lane]);
                if (a.flags) a.flags[qi] = (uint8_t)fl;
            }
            {
                const bool act = (active >> lane) & 1u;
                const uint32_t fl = act ? S.flags[lane] : 0u;
                w_ties += __popc(__ballot_sync(full, (fl & a.flag_tie) != 0));
                w_exact += __popc(__ballot_sync(full, (fl & a.flag_exact) != 0));
            }
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.cursor, 32ull);
            base = __shfl_sync(full, base, 0);
            if (base >= total) break;
            const unsigned long long pos = base + lane;
            const bool valid = pos < total;
            Vec3 o = {0.0, 0.0, 0.0};
            float4 q = make_float4(0.0f, 0.0f, 0.0f, INFINITY);
            uint32_t qi = 0;
            if (valid) {
                qi = a.perm ? a.perm[pos] : (uint32_t)pos;
                const double* p = a.pts + 3 * (uint64_t)qi;
                o = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
                q.x = (float)o.x; q.y = (float)o.y; q.z = (float)o.z;
            }
            const double ex = o.x - (double)q.x, ey = o.y - (double)q.y, ez = o.z - (double)q.z;   // exact differences
            const float perr = float_above(sqrt((ex * ex + ey * ey) + ez * ez) * 1.0000001) + 1e-37f;
            const double delta = 4e-16 * fmax(a.mag, fmax3(fabs(o.x), fabs(o.y), fabs(o.z)));
            const float dslack = float_above(2.0 * delta) + 1e-37f;
            // Walk down to the nearest leaf, nearest child first, pushing the far children that can still matter: every box holds a
            // primitive, so the distance to the FARTHEST corner of any box seen so far bounds the winner's distance from above
            // (far2, rounded up); a far child whose box is certainly beyond that (both sides moved by perr, the winner's side by the
            // rounding allowance of candidates) is dropped right away instead of being stacked and popped stale later.
            uint32_t cur = 0, firstleaf = kNone;
            float far2 = INFINITY;
            bool act = valid;
            while (__any_sync(full, act)) {
                if (ncnt + 32 > kWpNodeCap) {
                    if (!spill_bottom()) { if (lane == 0) atomicOr(&a.counters[2], 1ull); return; }
                }
                uint32_t farref = 0; float fard = 0.0f; bool stop = false, push = false;
                B2_WP_STAT(w_nodes += __popc(__ballot_sync(full, act)));
                if (act) {
                    const NodeView nd = load_node(a.nodes, cur);
                    float d0 = wp_box_sqdist(nd.lo0, nd.hi0, q), d1 = wp_box_sqdist(nd.lo1, nd.hi1, q);
                    far2 = fminf(far2, fminf(wp_box_far_sqdist(nd.lo0, nd.hi0, q), wp_box_far_sqdist(nd.lo1, nd.hi1, q)));
                    uint32_t c0 = nd.c0, c1 = nd.c1;
                    if (d1 < d0) { const uint32_t tc = c0; c0 = c1; c1 = tc; const float td = d0; d0 = d1; d1 = td; }
                    farref = c1; fard = d1;
                    const float reach = __fadd_ru(__fadd_ru(__fsqrt_ru(far2), __fmul_ru(2.0f, perr)), dslack);
                    push = fard <= __fmul_ru(reach, reach);
                    if (c0 & kLeafBit) { firstleaf = c0; stop = true; } else cur = c0;
                }
                const unsigned pm = __ballot_sync(full, push);
                S.npool[push ? ncnt + __popc(pm & lt) : kWpNodeCap] = make_uint2(farref, (__float_as_uint(fard) & ~31u) | lane);
                ncnt += __popc(pm);
                if (stop) act = false;
            }
            if (valid) {
                uint32_t id;
                const Tri3 t = load_tri(a.tris, firstleaf & ~kLeafBit, id);
                const CandOut co = wp_candidate<SEG>(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, o.x, o.y, o.z);
                const int nex0 = co.nexact;
                const Vec3 cpt = {co.x, co.y, co.z};
                const Vec3 v = cpt - o;
                const double sqd = (v.x * v.x + v.y * v.y) + v.z * v.z;
                __stcg(&G.o[0][lane], o.x); __stcg(&G.o[1][lane], o.y); __stcg(&G.o[2][lane], o.z);
                S.qa[lane] = q;
                S.qc[lane] = make_float4(INFINITY, __fadd_ru(__fmul_ru(1.7320509f, ev), perr), dslack, perr);
                __stcg(&G.bp[0][lane], cpt.x); __stcg(&G.bp[1][lane], cpt.y); __stcg(&G.bp[2][lane], cpt.z);
                __stcg(&G.bsqd[lane], sqd);
                S.bprim[lane] = id;
                wp_set_bounds(S, lane, sqd, v, delta, true);
                S.qidx[lane] = qi;
                S.flags[lane] = nex0 ? a.flag_exact : 0u;
            }
            active = __ballot_sync(full, valid);
            B2_WP_STAT(w_leaves += __popc(active); w_evals += __popc(active));
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------------------------------------ one round: every stage that is due, in order
        // A stage runs when its pool holds a full group (or nothing upstream can add to it) and the pool it feeds has room; the
        // stages are chained within a round, so a pair can travel from the node pool to the exact pool without the loop's housekeeping
        // in between (there are ~16 rounds per query).  Progress: if the node phase is held (lcnt >= 32) then pre-test 1 is due unless
        // mcnt >= 32, then pre-test 2 is due unless the exact pool is full, and a full exact pool (> kWpExactCap - 32 >= exact_go) is due itself;
        // with every upstream pool empty each stage drains.  (A per-round watchdog counter was measured at 8 % of the kernel's time and
        // is not kept; the GPU tests run every scheduling setting under a timeout instead.)
        if (ncnt > 0 && lcnt < 32) {
            // -------------------------------------------------------------------------------------------- (A) node phase
            do {
                if (ncnt + 32 > kWpNodeCap) {
                    if (!spill_bottom()) { if (lane == 0) atomicOr(&a.counters[2], 1ull); return; }
                }
                const int take = ncnt < 32 ? ncnt : 32;
                const bool has = (int)lane < take;
                ncnt -= take;
                const uint2 e = S.npool[has ? ncnt + (int)lane : kWpNodeCap];
                const unsigned owner = e.y & 31u;
                const float4 q = S.qa[owner];
                const bool live = has && __uint_as_float(e.y & ~31u) <= q.w;
                __syncwarp();                                      // all pops read before any push overwrites them
                const bool popped_leaf = live && (e.x & kLeafBit);
                const bool expand = live && !(e.x & kLeafBit);
                B2_WP_STAT(w_nodes += __popc(__ballot_sync(full, expand)));
                bool in0 = false, in1 = false;                     // 0 = nearer child, 1 = farther child
                uint32_t r0 = 0, r1 = 0;
                float d0 = 0.0f, d1 = 0.0f;
                // (Measured and rejected, profiles/r2_03: lane PAIRS fetching the two halves of one node with one load instruction, each lane testing the
                //  half it fetched of both nodes and shuffling the partner's results back: L1 wavefronts -5 %, instructions +16 %, 26.0 ms instead of
                //  24.0.  profiles/r2_04: 32-byte nodes with 16-bit grid boxes, one sector per visit: the L1 data stage drops from 74 % to 55 % of its peak,
                //  but decoding + 4 % more node and 9 % more leaf visits (boxes loosened by up to two cells) cost +17 % instructions: 25.6 ms.  The kernel
                //  is bound by instruction issue at the 20 warps per SM its registers and shared memory allow, not by the L1 wavefronts alone.)
                if (expand) {
                    const NodeView nd = load_node(a.nodes, e.x);
                    d0 = wp_box_sqdist(nd.lo0, nd.hi0, q); d1 = wp_box_sqdist(nd.lo1, nd.hi1, q);
                    r0 = nd.c0; r1 = nd.c1;
                    if (d1 < d0) { const uint32_t tc = r0; r0 = r1; r1 = tc; const float td = d0; d0 = d1; d1 = td; }
                    in0 = d0 <= q.w; in1 = d1 <= q.w;
                }
                // leaf pairs: a popped leaf entry or the nearer child, then the farther child (masked lanes write the sink entry)
                {
                    const bool l0 = popped_leaf || (in0 && (r0 & kLeafBit));
                    const uint32_t ref0 = popped_leaf ? e.x : r0;
                    const unsigned m0 = __ballot_sync(full, l0);
                    S.lpool[l0 ? lcnt + __popc(m0 & lt) : kWpLeafCap] = (owner << 27) | (ref0 & kWpRefMask);
                    lcnt += __popc(m0);
                    const bool l1 = in1 && (r1 & kLeafBit);
                    const unsigned m1 = __ballot_sync(full, l1);
                    S.lpool[l1 ? lcnt + __popc(m1 & lt) : kWpLeafCap] = (owner << 27) | (r1 & kWpRefMask);
                    lcnt += __popc(m1);
                }
                // internal children: farther first, nearer on top
                {
                    const bool p1 = in1 && !(r1 & kLeafBit);
                    const unsigned m1 = __ballot_sync(full, p1);
                    S.npool[p1 ? ncnt + __popc(m1 & lt) : kWpNodeCap] = make_uint2(r1, (__float_as_uint(d1) & ~31u) | owner);
                    ncnt += __popc(m1);
                    const bool p0 = in0 && !(r0 & kLeafBit);
                    const unsigned m0 = __ballot_sync(full, p0);
                    S.npool[p0 ? ncnt + __popc(m0 & lt) : kWpNodeCap] = make_uint2(r0, (__float_as_uint(d0) & ~31u) | owner);
                    ncnt += __popc(m0);
                }
                __syncwarp();
            } while (ncnt > 0 && lcnt < 32);
        }
        __syncwarp();                                          // the node phase's pool writes are visible to every lane
        if (mcnt < 32) {
            // -------------------------------------------------------------------------------------------- (B1) pre-test 1: two cheap directions
            // full groups of 32 from the top; what is left (< 32) waits for more unless the node pool has run dry
            while (lcnt >= 32 || (lcnt > 0 && ncnt == 0)) {
                const int take = lcnt < 32 ? lcnt : 32;
                lcnt -= take;
                bool survive = false;
                uint32_t pair = 0;
                float lb = 0.0f;
                if ((int)lane < take) {
                    pair = S.lpool[lcnt + (int)lane];
                    const unsigned owner = pair >> 27;
                    const RelTri r = wp_relative(a.trisf, pair & kWpRefMask, S.qa[owner]);
                    const float4 qc = S.qc[owner];
                    const float g[3] = {(r.ra[0] + r.rb[0]) + r.rc[0], (r.ra[1] + r.rb[1]) + r.rc[1], (r.ra[2] + r.rb[2]) + r.rc[2]};
                    // direction of the centroid: nearly optimal for a triangle that is small compared with its distance (the tangent plane of
                    // the slot's ball as a second direction culled nothing more and cost 3 %)
                    lb = wp_lower(g, sqrtf(dot3f(g, g)), r) - qc.y;
                    survive = !(lb > qc.x);
                }
                B2_WP_STAT(w_leaves += take);
                const unsigned sm = __ballot_sync(full, survive);
                if (survive) S.mpool[mcnt + __popc(sm & lt)] = make_uint2(pair, __float_as_uint(lb));
                mcnt += __popc(sm);
            }
        }
        __syncwarp();
        {
            // -------------------------------------------------------------------------------------------- (B2) pre-test 2: approximate closest point
            // full groups of 32 from the top (the rest waits unless nothing upstream is left); what does not fit waits for the exact phase
            while ((mcnt >= 32 || (mcnt > 0 && ncnt == 0 && lcnt == 0)) && ecnt + 32 <= kWpExactCap) {
                const int take = mcnt < 32 ? mcnt : 32;
                mcnt -= take;
                bool survive = false;
                uint2 e = make_uint2(0, 0);
                if ((int)lane < take) {
                    e = S.mpool[mcnt + (int)lane];
                    const unsigned owner = e.x >> 27;
                    const float4 qc = S.qc[owner];
                    if (!(__uint_as_float(e.y) > qc.x)) {           // (else: the ball has shrunk since pre-test 1)
                        const RelTri r = wp_relative(a.trisf, e.x & kWpRefMask, S.qa[owner]);
                        float lb, hi;
                        wp_sharp(r, qc.z, qc.y, &lb, &hi);
                        if (hi < qc.x) wp_shrink(S, owner, hi);   // this primitive is certainly nearer than the slot's ball: shrink it now, in binary32
                        lb = fmaxf(lb, __uint_as_float(e.y));
                        if (!(lb > qc.x)) { survive = true; e.y = __float_as_uint(lb); }
                    }
                }
                const unsigned sm = __ballot_sync(full, survive);
                if (survive) S.epool[ecnt + __popc(sm & lt)] = e;
                ecnt += __popc(sm);
            }
        }
        __syncwarp();
        if (ecnt >= a.exact_go || (ecnt > 0 && ncnt == 0 && lcnt == 0 && mcnt == 0 && spilled == 0)) {
            // -------------------------------------------------------------------------------------------- (B3) binary64 projection + exact comparison
            // (i) pairs whose lower bound has fallen outside their slot's ball retire without any binary64 work
            int kept = 0;
            for (int base = 0; base < ecnt; base += 32) {
                const int i = base + (int)lane;
                bool keep = false;
                uint2 e = make_uint2(0, 0);
                if (i < ecnt) {
                    e = S.epool[i];
                    const unsigned owner = e.x >> 27;
                    keep = !(__uint_as_float(e.y) > S.qc[owner].x);
                }
                __syncwarp();
                const unsigned km = __ballot_sync(full, keep);
                if (keep) S.epool[kept + __popc(km & lt)] = e;
                kept += __popc(km);
                __syncwarp();
            }
            ecnt = kept;
            const bool drain = ncnt == 0 && lcnt == 0 && mcnt == 0 && spilled == 0;
            if (ecnt >= a.exact_go || drain) {
                for (int base = 0; base < ecnt; base += 32) {
                    const int i = base + (int)lane;
                    const bool has = i < ecnt;
                    bool pending = false;
                    unsigned owner = 0;
                    uint32_t id = 0, key = 0;
                    Vec3 o = {0.0, 0.0, 0.0}, cpt = {0.0, 0.0, 0.0}, v = {0.0, 0.0, 0.0};
                    double sqd = 0.0;
                    if (has) {
                        const uint2 e = S.epool[i];
                        owner = e.x >> 27;
                        if (!(__uint_as_float(e.y) > S.qc[owner].x)) {   // (else: an earlier batch of this very phase has shrunk the ball)
                            const Tri3 t = load_tri(a.tris, e.x & kWpRefMask, id);
                            o = {__ldcg(&G.o[0][owner]), __ldcg(&G.o[1][owner]), __ldcg(&G.o[2][owner])};
                            const CandOut co = wp_candidate<SEG>(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, o.x, o.y, o.z);
                            cpt = {co.x, co.y, co.z};
                            if (co.nexact) atomicOr(&S.flags[owner], a.flag_exact);
                            v = cpt - o;
                            sqd = (v.x * v.x + v.y * v.y) + v.z * v.z;   // Compute_squared_distance_3(query, closest): AABB_tree.h:1024-1032
                            key = __float_as_uint((float)sqd);
                            pending = true;
                        }
                    }
                    // Every waiting candidate is compared (exactly) with its slot's incumbent, all lanes at once; the ones that are farther
                    // retire.  Of the rest, ONE per slot (the apparently nearest) replaces the incumbent, the others meet the new incumbent in
                    // the next round.  Each replacement is a step the sequential reference loop could have taken, and a candidate farther
                    // than an earlier incumbent is farther than every later one, so the outcome is that of the sequential loop.
                    unsigned pm = __ballot_sync(full, pending);
                    B2_WP_STAT(w_evals += __popc(pm));
                    while (pm) {
                        int sg = 0;
                        if (pending) {
                            const Vec3 bp = {__ldcg(&G.bp[0][owner]), __ldcg(&G.bp[1][owner]), __ldcg(&G.bp[2][owner])};
                            int nexact = 0;
                            sg = cmp_dist(o, cpt, bp, nexact);
                            if (nexact) atomicOr(&S.flags[owner], a.flag_exact);
                            if (sg > 0) pending = false;
                        }
                        __syncwarp();                              // incumbents read before any is replaced
                        const unsigned cm = __ballot_sync(full, pending);
                        if (pending) {
                            const unsigned grp = __match_any_sync(cm, owner);
                            const unsigned gmin = __reduce_min_sync(grp, key);
                            const unsigned wm = __ballot_sync(cm, key == gmin) & grp;
                            if ((unsigned)(__ffs(wm) - 1) == lane) { wp_apply(S, G, owner, sg, cpt, v, sqd, 4e-16 * fmax(a.mag, fmax3(fabs(o.x), fabs(o.y), fabs(o.z))), id, a.flag_tie); pending = false; }
                        }
                        __syncwarp();
                        pm = __ballot_sync(full, pending);
                    }
                }
                ecnt = 0;
            }
        }
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(&a.counters[3], (unsigned long long)w_nodes);
        atomicAdd(&a.counters[4], (unsigned long long)w_leaves);
        atomicAdd(&a.counters[6], (unsigned long long)(w_leaves - w_evals));   // leaves that never reached the binary64 projection
        if (w_exact) atomicAdd(&a.counters[0], (unsigned long long)w_exact);
        if (w_ties) atomicAdd(&a.counters[1], (unsigned long long)w_ties);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// closest_point in TWO PASSES (B200AABB_OPT_CLOSEST_KERNEL = 3).
//   closest_filter_kernel   one thread per query, binary32 only: the traversal of the pool kernel's stages (A), (B1), (B2) — same rigorous bounds, wp_box_sqdist /
//                           wp_lower / wp_sharp — with a private stack and NO binary64 work: the ball shrinks with the rigorous upper bounds alone, and every
//                           primitive whose lower bound stays inside the final ball is a CANDIDATE (at most kFcCand kept; more = the query is handed to the pool kernel)
//   closest_resolve_kernel  one thread per query: the reference's projection on each candidate (binary64) and the exact comparison; winner = smallest exact distance,
//                           ties -> smallest primitive id + FLAG_TIE, as everywhere
// Every primitive that attains the minimum has a true distance <= the final radius, so its lower bound is inside the ball at every moment: it is a candidate.
// Why: the pool kernel spends ~180 thread instructions per (query, node / leaf) pair, most of them warp-wide bookkeeping of its pools (profiles/r2_17); a private
// traversal in binary32 needs ~50, and without the binary64 stage in the loop its warps no longer diverge between cheap and expensive phases (what made the
// thread-per-query kernel of round 1 run with 5 of 32 lanes, profiles/r1_03).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFcCand = 8;
constexpr uint8_t kFcOverflow = 255;
struct FilterClosestArgs {
    const BvhNode* nodes; const TriRec* tris; const TriF* trisf; uint32_t root; double mag;
    const uint32_t* perm; unsigned long long* counters; const double* pts; uint64_t n;
    uint32_t* cand;                // n x kFcCand leaf slots
    uint8_t* ncand;                // per query: candidates kept, or kFcOverflow
    uint32_t* overflow;            // queries handed to the pool kernel
    unsigned long long* overflow_count;
    double* cp; uint32_t* prim; double* sqd; uint8_t* flags;
    uint32_t flag_tie, flag_exact;
};

__global__ void __launch_bounds__(128, 6) closest_filter_kernel(const FilterClosestArgs a) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v_nodes = 0, v_leaves = 0, v_culled = 0;
    if (slot < a.n) {
        const uint64_t qi = a.perm ? a.perm[slot] : slot;
        const double* p = a.pts + 3 * qi;
        const Vec3 o = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        float4 q = make_float4((float)o.x, (float)o.y, (float)o.z, INFINITY);
        const double ex = o.x - (double)q.x, ey = o.y - (double)q.y, ez = o.z - (double)q.z;
        const float perr = float_above(sqrt((ex * ex + ey * ey) + ez * ez) * 1.0000001) + 1e-37f;
        const double delta = 4e-16 * fmax(a.mag, fmax3(fabs(o.x), fabs(o.y), fabs(o.z)));
        const float dslack = float_above(2.0 * delta) + 1e-37f;
        const float ev = float_above(a.mag * 5.9604645e-8 * 1.000001);
        const float e2 = __fadd_ru(__fmul_ru(1.7320509f, ev), perr);
        float rbound = INFINITY;
        uint32_t cl[kFcCand];
        float clb[kFcCand];
        int nc = 0;
        bool ovf = false;
        uint32_t sref[kStackCap];
        float sd[kStackCap];
        int sp = 0, np = 0;
        uint32_t pend[4];
        uint32_t cur = a.root;
        if (cur & kLeafBit) { pend[np++] = cur; cur = kNone; }
        while (true) {
            while (cur != kNone && np < 3) {                             // node phase: stage (A) of the pool kernel
                const NodeView nd = load_node(a.nodes, cur);
                ++v_nodes;
                float d0 = wp_box_sqdist(nd.lo0, nd.hi0, q), d1 = wp_box_sqdist(nd.lo1, nd.hi1, q);
                uint32_t c0 = nd.c0, c1 = nd.c1;
                if (d1 < d0) { const uint32_t tc = c0; c0 = c1; c1 = tc; const float td = d0; d0 = d1; d1 = td; }
                uint32_t next = kNone;
                if (d0 <= q.w) { if (c0 & kLeafBit) pend[np++] = c0; else next = c0; }
                if (d1 <= q.w) {
                    if (c1 & kLeafBit) pend[np++] = c1;
                    else if (next == kNone) next = c1;
                    else if (sp < kStackCap) { sref[sp] = c1; sd[sp] = d1; ++sp; }
                    else ovf = true;                                     // (deeper than any tree this library builds: the pool kernel answers the query)
                }
                while (next == kNone && sp > 0) {
                    --sp;
                    if (sd[sp] <= q.w) next = sref[sp];
                }
                cur = next;
            }
            for (int k = 0; k < np; ++k) {                               // leaf phase: stages (B1) and (B2)
                const uint32_t leaf = pend[k] & ~kLeafBit;
                ++v_leaves;
                const RelTri r = wp_relative(a.trisf, leaf, q);
                const float g[3] = {(r.ra[0] + r.rb[0]) + r.rc[0], (r.ra[1] + r.rb[1]) + r.rc[1], (r.ra[2] + r.rb[2]) + r.rc[2]};
                const float lb1 = wp_lower(g, sqrtf(dot3f(g, g)), r) - e2;
                if (lb1 > rbound) { ++v_culled; continue; }
                float lb, hi;
                wp_sharp(r, dslack, e2, &lb, &hi);
                if (hi < rbound) { rbound = hi; const float rr = __fadd_ru(rbound, perr); q.w = __fmul_ru(rr, rr); }
                lb = fmaxf(lb, lb1);
                if (lb > rbound) { ++v_culled; continue; }
                if (nc == kFcCand) {                                     // full: drop the candidates the ball has left behind
                    int m = 0;
                    for (int j = 0; j < kFcCand; ++j) if (!(clb[j] > rbound)) { cl[m] = cl[j]; clb[m] = clb[j]; ++m; }
                    nc = m;
                }
                if (nc == kFcCand) ovf = true; else { cl[nc] = leaf; clb[nc] = lb; ++nc; }
            }
            np = 0;
            if (ovf) break;
            if (cur == kNone) {
                while (cur == kNone && sp > 0) {
                    --sp;
                    if (sd[sp] <= q.w) cur = sref[sp];
                }
                if (cur == kNone) break;
            }
        }
        if (ovf) {
            a.ncand[qi] = kFcOverflow;
            a.overflow[atomicAdd(a.overflow_count, 1ull)] = (uint32_t)qi;
        } else {
            int m = 0;
            for (int j = 0; j < nc; ++j) if (!(clb[j] > rbound)) a.cand[qi * kFcCand + m++] = cl[j];
            a.ncand[qi] = (uint8_t)m;
        }
    }
    const unsigned full = 0xFFFFFFFFu;
    const unsigned nodes = __reduce_add_sync(full, v_nodes), leaves = __reduce_add_sync(full, v_leaves), culled = __reduce_add_sync(full, v_culled);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&a.counters[3], (unsigned long long)nodes);
        atomicAdd(&a.counters[4], (unsigned long long)leaves);
        atomicAdd(&a.counters[6], (unsigned long long)culled);
    }
}

__global__ void __launch_bounds__(128) closest_resolve_kernel(const FilterClosestArgs a) {
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool tie = false, exact = false;
    if (slot < a.n) {
        const uint64_t qi = a.perm ? a.perm[slot] : slot;
        const int nc = a.ncand[qi];
        if (nc != kFcOverflow) {
            const double* p = a.pts + 3 * qi;
            const Vec3 o = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
            Vec3 bp = o;
            double bsqd = 0.0;
            uint32_t bid = kNone;
            for (int j = 0; j < nc; ++j) {
                uint32_t id;
                const Tri3 t = load_tri(a.tris, a.cand[qi * kFcCand + j], id);
                const CandOut co = wp_candidate(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, o.x, o.y, o.z);
                exact |= co.nexact != 0;
                const Vec3 cpt = {co.x, co.y, co.z};
                bool take = bid == kNone;
                if (!take) {
                    int nexact = 0;
                    const int sg = cmp_dist(o, cpt, bp, nexact);
                    exact |= nexact != 0;
                    if (sg == 0) { tie = true; take = id < bid; }
                    else if (sg < 0) { tie = false; take = true; }
                }
                if (take) {
                    const Vec3 v = cpt - o;
                    bp = cpt; bid = id; bsqd = (v.x * v.x + v.y * v.y) + v.z * v.z;   // Compute_squared_distance_3(query, closest): AABB_tree.h:1024-1032
                }
            }
            if (a.cp) { a.cp[3 * qi] = bp.x; a.cp[3 * qi + 1] = bp.y; a.cp[3 * qi + 2] = bp.z; }
            if (a.prim) a.prim[qi] = bid;
            if (a.sqd) a.sqd[qi] = bsqd;
            if (a.flags) a.flags[qi] = (uint8_t)((tie ? a.flag_tie : 0u) | (exact ? a.flag_exact : 0u));
        } else { tie = false; exact = false; }
    }
    const unsigned full = 0xFFFFFFFFu;
    const unsigned ties = __reduce_add_sync(full, tie ? 1u : 0u), exacts = __reduce_add_sync(full, exact ? 1u : 0u);
    if ((threadIdx.x & 31) == 0) {
        if (ties) atomicAdd(&a.counters[1], (unsigned long long)ties);
        if (exacts) atomicAdd(&a.counters[0], (unsigned long long)exacts);
    }
}

#endif  // __CUDACC__

}  // namespace b2
