// bvh.cuh — HBM layout of the tree and the per-query traversal routines.
//
// Layout:
//   BvhNode   64 B  one INTERNAL node of the binary LBVH.  It carries the boxes of BOTH children as binary32 intervals
//                   rounded OUTWARD from the exact binary64 boxes (lo towards -inf, hi towards +inf), so a traversal step
//                   costs two 32-byte sectors and leaf boxes never need their own node.  N-1 nodes for N primitives.
//   TriRec    96 B  one primitive in Morton order: 9 binary64 coordinates (v0,v1,v2), the primitive id, padding to 3 sectors.
//   TriF      64 B  the same primitive rounded to binary32 (pre-tests of closest_warp.cuh only; never decides a result).
// Child reference: bit 31 set = leaf, low bits = slot in the TriRec array; else index of an internal node.
//
// Traversal replaces (paths relative to the CGAL source root)
//   AABB_tree/include/CGAL/AABB_tree/internal/AABB_node.h:129-167              recursive traversal
//   AABB_tree/include/CGAL/AABB_tree/internal/AABB_traversal_traits.h:150-327  Listing/Do_intersect/Projection traits
//   AABB_tree/include/CGAL/AABB_tree/internal/AABB_ray_intersection.h:49-150   best-first first_intersection
// by an iterative stack traversal per query.  Node tests are CONSERVATIVE binary32 tests with explicit error terms
// (they never reject a box that contains a qualifying primitive) and run on the FP32 pipe; every decision that reaches
// the result is taken by the exact binary64 leaf routines of geom.cuh, so booleans, counts and winners do not depend on
// the tree shape or on the box precision (SURVEY.md §3.4, §7 hard part 1).
#pragma once
#include <stdint.h>
#include "geom.cuh"


namespace b2 {

constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kStackCap = 96;

// Each 32-byte half of a node is self-contained (one child's box + its reference): a thread that needs both children reads the two
// sectors with two 256-bit loads, and the lane PAIRS of closest_pair.cuh read one half each, so a node costs ONE L1 wavefront per pair.
struct __align__(32) BvhNode {
    float lo0[3], hi0[3];  // child 0
    uint32_t child0;
    uint32_t parent;       // build / refit only
    float lo1[3], hi1[3];  // child 1
    uint32_t child1;
    uint32_t visit;        // refit: the other end of the node's leaf range, then the arrival flag (build only)
    B2_HD uint32_t child(int c) const { return c ? child1 : child0; }
};
static_assert(sizeof(BvhNode) == 64, "node is two 32-byte sectors");

struct __align__(32) TriRec {
    double v[9];
    uint32_t id;
    uint32_t pad[5];
};
static_assert(sizeof(TriRec) == 96, "triangle record is three 32-byte sectors");

// binary32 copy of a primitive (same order as the TriRec array): each coordinate rounded to nearest.  Read by the rigorous binary32
// pre-tests of the warp-cooperative closest_point kernel (closest_warp.cuh), which account for the rounding explicitly; 64 bytes
// = two 256-bit loads (two L1 wavefronts per lane) instead of the six 128-bit ones of a TriRec.
struct __align__(32) TriF {
    float v[9];
    uint32_t pad[7];
};
static_assert(sizeof(TriF) == 64, "binary32 triangle record is two 32-byte sectors");

// per-query traversal statistics (accumulated into the handle's counters by the kernels)
#ifndef B2_TRACE_NODE
#define B2_TRACE_NODE(index)          // tests/hostcheck defines it to record which nodes traverse_closest fetches (depth histogram)
#endif
struct TravStat { int nexact = 0; int overflow = 0; uint32_t nodes = 0, leaves = 0, culled = 0; };   // culled: leaves rejected by the binary32 pre-test

// directed roundings binary64 -> binary32 (any finite input; overflow saturates to +-FLT_MAX / +-inf on the safe side)
B2_HD float float_below(double v) {
#if defined(__CUDA_ARCH__)
    return __double2float_rd(v);
#else
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return f;
#endif
}
B2_HD float float_above(double v) {
#if defined(__CUDA_ARCH__)
    return __double2float_ru(v);
#else
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
#endif
}
// directed binary32 arithmetic (single instructions with a rounding modifier on the device)
B2_HD float sub_dn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rd(a, b);
#else
    return float_below((double)a - (double)b);      // the binary64 difference of two floats is exact
#endif
}
B2_HD float mul_dn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rd(a, b);
#else
    return float_below((double)a * (double)b);      // the binary64 product of two floats is exact
#endif
}
B2_HD float add_dn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rd(a, b);
#else
    return float_below((double)a + (double)b);
#endif
}

struct NodeView { float lo0[3], hi0[3], lo1[3], hi1[3]; uint32_t c0, c1; };

#if defined(__CUDACC__)
// 256-bit read-only load (sm_100: LDG.E.256): one request per 32-byte sector and lane instead of two 128-bit ones, which halves the
// L1 wavefronts of the scattered node / triangle fetches that bound the traversal kernels (profiles/r1_05); measured against two
// 128-bit loads on the final kernels: closest_point 290 vs 282 M queries/s, rays equal.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
#endif

B2_HD NodeView load_node(const BvhNode* __restrict__ nodes, uint32_t i) {
    NodeView n;
#if defined(__CUDA_ARCH__)
    uint32_t a[8], b[8];
    ldg256(nodes + i, a);
    ldg256(reinterpret_cast<const char*>(nodes + i) + 32, b);
    n.lo0[0] = __uint_as_float(a[0]); n.lo0[1] = __uint_as_float(a[1]); n.lo0[2] = __uint_as_float(a[2]);
    n.hi0[0] = __uint_as_float(a[3]); n.hi0[1] = __uint_as_float(a[4]); n.hi0[2] = __uint_as_float(a[5]);
    n.lo1[0] = __uint_as_float(b[0]); n.lo1[1] = __uint_as_float(b[1]); n.lo1[2] = __uint_as_float(b[2]);
    n.hi1[0] = __uint_as_float(b[3]); n.hi1[1] = __uint_as_float(b[4]); n.hi1[2] = __uint_as_float(b[5]);
    n.c0 = a[6]; n.c1 = b[6];
#else
    const BvhNode& s = nodes[i];
    for (int k = 0; k < 3; ++k) { n.lo0[k] = s.lo0[k]; n.hi0[k] = s.hi0[k]; n.lo1[k] = s.lo1[k]; n.hi1[k] = s.hi1[k]; }
    n.c0 = s.child0; n.c1 = s.child1;
#endif
    return n;
}

B2_HD Tri3 load_tri(const TriRec* __restrict__ tris, uint32_t slot, uint32_t& id) {
    Tri3 t;
#if defined(__CUDA_ARCH__)
    uint32_t a[8], b[8], c[8];
    const char* p = reinterpret_cast<const char*>(tris + slot);
    ldg256(p, a); ldg256(p + 32, b); ldg256(p + 64, c);
    auto D = [](uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); };
    t.a = {D(a[0], a[1]), D(a[2], a[3]), D(a[4], a[5])};
    t.b = {D(a[6], a[7]), D(b[0], b[1]), D(b[2], b[3])};
    t.c = {D(b[4], b[5]), D(b[6], b[7]), D(c[0], c[1])};
    id = c[2];
#else
    const TriRec& s = tris[slot];
    t.a = {s.v[0], s.v[1], s.v[2]}; t.b = {s.v[3], s.v[4], s.v[5]}; t.c = {s.v[6], s.v[7], s.v[8]};
    id = s.id;
#endif
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// Conservative node tests (binary32 with explicit error terms)
// ---------------------------------------------------------------------------------------------------------------
struct LineCtx {              // Segment_3 / Ray_3 query, parametrised x(t) = s + t (q - s)
    Vec3 s, q, d;             // d = q - s (rounded once, binary64)
    float sf[3];              // s rounded to nearest binary32
    float inv[3];             // 1/d per axis rounded to binary32 (unused where d == 0)
    float err[3];             // |s - sf| * |1/d| (rounded up): error of a slab parameter caused by rounding the source
    int zero_mask;            // axes with d == 0
    int major;                // argmax |d| (first max wins), AABB_ray_intersection.h:171-176
    double t_abs_slack;       // absolute error allowance on computed hit parameters
};
// `mag` = largest coordinate magnitude of the scene (root box); hit points carry a few ulp(mag) of rounding error
B2_HD LineCtx make_line(const Vec3& s, const Vec3& q, double mag = 0.0) {
    LineCtx c; c.s = s; c.q = q; c.d = q - s;
    c.zero_mask = (c.d.x == 0.0 ? 1 : 0) | (c.d.y == 0.0 ? 2 : 0) | (c.d.z == 0.0 ? 4 : 0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double si = comp(s, i), inv = 1.0 / comp(c.d, i);
        c.sf[i] = (float)si;
        c.inv[i] = (float)inv;
        c.err[i] = float_above(fabs(si - (double)c.sf[i]) * fabs(inv) * 1.0000001) + 1e-37f;   // + underflow guard
    }
    c.major = ray_major_axis(c.d);
    const double smax = fmax(mag, fmax3(fabs(s.x), fabs(s.y), fabs(s.z)));
    c.t_abs_slack = 1.5e-14 * smax / fabs(comp(c.d, c.major));
    return c;
}
// the same for a Line_3 given as point + direction vector (x(t) = p + t v, t unbounded)
B2_HD LineCtx make_line_pv(const Vec3& p, const Vec3& v) {
    LineCtx c = make_line(p, p, 0.0);
    c.q = p + v; c.d = v;
    c.zero_mask = (v.x == 0.0 ? 1 : 0) | (v.y == 0.0 ? 2 : 0) | (v.z == 0.0 ? 4 : 0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double si = comp(p, i), inv = 1.0 / comp(v, i);
        c.inv[i] = (float)inv;
        c.err[i] = float_above(fabs(si - (double)c.sf[i]) * fabs(inv) * 1.0000001) + 1e-37f;
    }
    c.major = ray_major_axis(v);
    return c;
}
// Does x(t), t in [0, tlimit], possibly meet the closed box [lo, hi]?  tn = lower bound of the entry parameter.
// Per axis the computed parameter t_c = (b - sf) * inv differs from the true (b - s) / d by at most
// |t_c| * 2^-21 (one subtraction, one product, the rounding of 1/d and of d itself: < 4 * 2^-24 relative) + err[i]
// (rounding of the source).  NaNs (0 * inf, inf - inf) are dropped by fminf/fmaxf, which only loosens the interval.
B2_HD bool line_box(const float* lo, const float* hi, const LineCtx& c, float tlimit, float& tn, float tstart = 0.0f) {
    float tnear = tstart, tfar = tlimit;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (c.zero_mask & (1 << i)) {
            const double si = comp(c.s, i);
            if (si < (double)lo[i] || si > (double)hi[i]) return false;
        } else {
            const float t1 = (lo[i] - c.sf[i]) * c.inv[i], t2 = (hi[i] - c.sf[i]) * c.inv[i];
            float a = fminf(t1, t2), b = fmaxf(t1, t2);
            a = a - (fabsf(a) * 4.76837158e-7f + c.err[i]);
            b = b + (fabsf(b) * 4.76837158e-7f + c.err[i]);
            tnear = fmaxf(tnear, a);
            tfar = fminf(tfar, b);
        }
    }
    tn = tnear;
    return tnear <= tfar;
}

struct PointCtx { float up[3], dn[3]; };   // the query point rounded up / down to binary32
B2_HD PointCtx make_point(const Vec3& p) {
    PointCtx c;
    c.up[0] = float_above(p.x); c.up[1] = float_above(p.y); c.up[2] = float_above(p.z);
    c.dn[0] = float_below(p.x); c.dn[1] = float_below(p.y); c.dn[2] = float_below(p.z);
    return c;
}
// LOWER bound of the squared distance from the query point to the closed box: every operation rounds towards the
// smaller result (lo - p >= lo - p_up, p - hi >= p_dn - hi; sums and products of non-negative terms rounded down)
B2_HD float point_box_sqdist(const float* lo, const float* hi, const PointCtx& c) {
    const float dx = fmaxf(fmaxf(sub_dn(lo[0], c.up[0]), sub_dn(c.dn[0], hi[0])), 0.0f);
    const float dy = fmaxf(fmaxf(sub_dn(lo[1], c.up[1]), sub_dn(c.dn[1], hi[1])), 0.0f);
    const float dz = fmaxf(fmaxf(sub_dn(lo[2], c.up[2]), sub_dn(c.dn[2], hi[2])), 0.0f);
    return add_dn(add_dn(mul_dn(dx, dx), mul_dn(dy, dy)), mul_dn(dz, dz));
}

// ---------------------------------------------------------------------------------------------------------------
// Traversal scheme shared by the three routines below ("while-while" with deferred leaves).  A thread alternates between
//   a NODE phase   — pops / descends internal nodes, tests the two child boxes in binary32, and only RECORDS the leaf
//                    children that pass (a few pending slots per thread), and
//   a LEAF phase   — runs the binary64 primitive routines of geom.cuh on the recorded leaves.
// Both phases are loops, so the 32 threads of a warp reconverge at each phase boundary: the long exact leaf routines are
// executed by all threads that have a pending leaf at once instead of by one thread at a time while the others wait
// (measured on the one-loop version, profiles/r1_01: 2.9 of 32 threads active per instruction in closest_point).
// ---------------------------------------------------------------------------------------------------------------
// ---- boxes of Triangle_3 queries and trees under an affine transformation (used by traverse_line / traverse_triangle below)
struct BoxCtx { float lo[3], hi[3]; };
B2_HD BoxCtx make_box_of(const Tri3& t) {
    BoxCtx c;
    c.lo[0] = float_below(fmin3(t.a.x, t.b.x, t.c.x)); c.lo[1] = float_below(fmin3(t.a.y, t.b.y, t.c.y)); c.lo[2] = float_below(fmin3(t.a.z, t.b.z, t.c.z));
    c.hi[0] = float_above(fmax3(t.a.x, t.b.x, t.c.x)); c.hi[1] = float_above(fmax3(t.a.y, t.b.y, t.c.y)); c.hi[2] = float_above(fmax3(t.a.z, t.b.z, t.c.z));
    return c;
}
B2_HD bool box_box(const float* lo, const float* hi, const BoxCtx& c) {
    return lo[0] <= c.hi[0] && c.lo[0] <= hi[0] && lo[1] <= c.hi[1] && c.lo[1] <= hi[1] && lo[2] <= c.hi[2] && c.lo[2] <= hi[2];
}
// The traversed tree may stand under an affine transformation (Rigid_triangle_mesh_collision_detection.h: each mesh keeps its tree, built once, and an
// Aff_transformation_3; Polygon_mesh_processing/internal/AABB_traversal_traits_with_transformation.h:141-156 transforms the DATUM of each primitive and the box of
// each node on the fly).  m = the 3 x 4 matrix (m(i, j) of Aff_transformation_3, row major).  A point is transformed as Aff_transformation_repC3::transform
// does (Cartesian/Aff_transformation_rep_3.h:111-117: t11 x + t12 y + t13 z + t14, left to right, no fused operations) — for the translation and scaling
// representations the same expression gives their results (1 x + 0 y + 0 z + t = x + t; s x + 0 + 0 + 0 = s x) — and the exact predicates see the rounded image.
struct NoXf {
    static constexpr bool kIdentity = true;
    B2_HD Tri3 apply(const Tri3& t) const { return t; }
    B2_HD bool box_meets(const float* lo, const float* hi, const BoxCtx& c) const { return box_box(lo, hi, c); }
};
struct AffXf {
    static constexpr bool kIdentity = false;
    double m[12];
    B2_HD Vec3 point(const Vec3& p) const {
        return {((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7], ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]};
    }
    B2_HD Tri3 apply(const Tri3& t) const { return {point(t.a), point(t.b), point(t.c)}; }
    // image of the box as centre +- |M| half-extent (the reference takes the box of the 8 transformed corners in interval arithmetic, :36-68: the same box
    // up to rounding); widened by a bound of its own rounding and of the rounding of the transformed primitives it has to contain
    // the same image as a binary32 box rounded outward (for the line / box tests of the ray traversals)
    B2_HD void image_box(const float* lo, const float* hi, float* olo, float* ohi) const {
        const double cx = 0.5 * ((double)lo[0] + (double)hi[0]), cy = 0.5 * ((double)lo[1] + (double)hi[1]), cz = 0.5 * ((double)lo[2] + (double)hi[2]);
        const double hx = 0.5 * ((double)hi[0] - (double)lo[0]), hy = 0.5 * ((double)hi[1] - (double)lo[1]), hz = 0.5 * ((double)hi[2] - (double)lo[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* r = m + 4 * k;
            const double ctr = ((r[0] * cx + r[1] * cy) + r[2] * cz) + r[3];
            const double ext = (fabs(r[0]) * hx + fabs(r[1]) * hy) + fabs(r[2]) * hz;
            const double mag = ((fabs(r[0]) * (fabs(cx) + hx) + fabs(r[1]) * (fabs(cy) + hy)) + fabs(r[2]) * (fabs(cz) + hz)) + fabs(r[3]);
            const double wide = ext + 1e-14 * mag;
            const bool ok = ctr == ctr && wide == wide;                       // (NaN: the whole axis)
            olo[k] = ok ? float_below(ctr - wide) : -INFINITY; ohi[k] = ok ? float_above(ctr + wide) : INFINITY;
        }
    }
    B2_HD bool box_meets(const float* lo, const float* hi, const BoxCtx& c) const {
        const double cx = 0.5 * ((double)lo[0] + (double)hi[0]), cy = 0.5 * ((double)lo[1] + (double)hi[1]), cz = 0.5 * ((double)lo[2] + (double)hi[2]);
        const double hx = 0.5 * ((double)hi[0] - (double)lo[0]), hy = 0.5 * ((double)hi[1] - (double)lo[1]), hz = 0.5 * ((double)hi[2] - (double)lo[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* r = m + 4 * k;
            const double ctr = ((r[0] * cx + r[1] * cy) + r[2] * cz) + r[3];
            const double ext = (fabs(r[0]) * hx + fabs(r[1]) * hy) + fabs(r[2]) * hz;
            const double mag = ((fabs(r[0]) * (fabs(cx) + hx) + fabs(r[1]) * (fabs(cy) + hy)) + fabs(r[2]) * (fabs(cz) + hz)) + fabs(r[3]);
            const double wide = ext + 1e-14 * mag;
            if (!(ctr - wide <= (double)c.hi[k] && (double)c.lo[k] <= ctr + wide)) return !(ctr == ctr && wide == wide);   // (NaN / overflow: keep the box)
        }
        return true;
    }
};
// warp vote used ONLY as a scheduling hint (when to run a deferred phase); the host emulation sees the thread alone
B2_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), p);
#else
    return p;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// do_intersect / number_of_intersected_primitives   (AABB_tree.h:903-932)
// ---------------------------------------------------------------------------------------------------------------
// MODE: kAnyHit   do_intersect (AABB_tree.h:903-913): 1 at the first intersected primitive, else 0
//       kCount    number_of_intersected_primitives (:915-932)
//       kAnyId    any_intersected_primitive (:947-956, First_primitive_traits): id of the first intersected primitive met, kNone if none
//       kList     all_intersected_primitives (:934-945, Listing_primitive_traits): ids written to `list`, returns how many
enum IntersectMode { kAnyHit = 0, kCount = 1, kAnyId = 2, kList = 3 };

// The traversal shared by all intersection queries: visits every leaf whose box the query line meets and calls leaf(ref) -> bool
// ("this primitive counts"); returns how many did.  STOP_AT_FIRST ends at the first one.
enum LineKind { kSegmentQ = 0, kRayQ = 1, kLineQ = 2 };
template <class XF>
B2_HD bool line_box_under(const XF& xf, const float* lo, const float* hi, const LineCtx& c, float tlimit, float& tn, float tstart) {
    if constexpr (XF::kIdentity) return line_box(lo, hi, c, tlimit, tn, tstart);
    else { float l[3], h[3]; xf.image_box(lo, hi, l, h); return line_box(l, h, c, tlimit, tn, tstart); }
}
template <int KIND, bool STOP_AT_FIRST, class Leaf, class XF = NoXf>
B2_HD uint32_t traverse_line(const BvhNode* __restrict__ nodes, uint32_t root, const LineCtx& c, TravStat& st, Leaf&& leaf, const XF& xf = XF()) {
    const float tlimit = KIND == kSegmentQ ? 1.0f : INFINITY, tstart = KIND == kLineQ ? -INFINITY : 0.0f;
    uint32_t found = 0;
    if (root & kLeafBit) return leaf(root) ? 1u : 0u;       // singleton tree: AABB_tree.h:516-525
    uint32_t stack[kStackCap];
    int sp = 0;
    // Leaves awaiting the exact test: a thread collects up to kPendCount before it leaves the node phase, so that the (long) exact
    // tests of a warp's threads run together.  Measured on the config 3 / 4 segment batches (scripts/tune_anyhit.py): 2.1x for the
    // counting modes, 1.2x for the modes that stop at the first hit although they may traverse a little further than needed.
    constexpr int kPendCount = 6, kPendGoAt = kPendCount - 1;
    uint32_t pend[kPendCount + 1];
    int np = 0;
    uint32_t cur = root;
    while (true) {
        while (cur != kNone && np < kPendGoAt) {          // node phase
            const NodeView n = load_node(nodes, cur);
            ++st.nodes;
            float tn;
            const bool h0 = line_box_under(xf, n.lo0, n.hi0, c, tlimit, tn, tstart);
            const bool h1 = line_box_under(xf, n.lo1, n.hi1, c, tlimit, tn, tstart);
            uint32_t next = kNone;
            if (h0) { if (n.c0 & kLeafBit) pend[np++] = n.c0; else next = n.c0; }
            if (h1) {
                if (n.c1 & kLeafBit) pend[np++] = n.c1;
                else if (next == kNone) next = n.c1;
                else if (sp < kStackCap) stack[sp++] = n.c1;
                else st.overflow = 1;
            }
            if (next == kNone && sp > 0) next = stack[--sp];
            cur = next;
        }
        while (np > 0) {                                  // leaf phase
            if (leaf(pend[--np])) { ++found; if (STOP_AT_FIRST) return found; }
        }
        if (cur == kNone) break;
    }
    return found;
}

// KIND: kSegmentQ (s, q = source, target), kRayQ (source, second point), kLineQ (s = a point of the line, q = its DIRECTION VECTOR)
template <int KIND, int MODE>
B2_HD uint32_t traverse_intersect(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root,
                                  const Vec3& s, const Vec3& q, TravStat& st, uint32_t* list = nullptr, uint32_t list_cap = 0) {
    const LineCtx c = KIND == kLineQ ? make_line_pv(s, q) : make_line(s, q);
    uint32_t nlisted = 0, found_id = kNone;
    constexpr bool kStopAtFirst = MODE == kAnyHit || MODE == kAnyId;
    const uint32_t found = traverse_line<KIND, kStopAtFirst>(nodes, root, c, st, [&](uint32_t ref) -> bool {
        uint32_t id;
        const Tri3 t = load_tri(tris, ref & ~kLeafBit, id);
        ++st.leaves;
        if (!(KIND == kRayQ ? ray_tri_do_intersect(t, s, q, st.nexact) : KIND == kLineQ ? line_tri_do_intersect(t, s, q, st.nexact) : seg_tri_do_intersect(t, s, q, st.nexact))) return false;
        if (MODE == kList) { if (nlisted < list_cap) list[nlisted] = id; ++nlisted; }   // (never writes past the caller's room; the count tells)
        if (MODE == kAnyId) found_id = id;
        return true;
    });
    return MODE == kAnyId ? found_id : (MODE == kAnyHit ? (found ? 1u : 0u) : found);
}

// Triangle_3 queries (do_intersect / number_of_intersected_primitives / any_intersected_primitive / all_intersected_primitives with a Triangle_3,
// AABB_tree.h:903-956; the query a tree-vs-tree do_intersect issues per primitive, AABB_traits_3.h:679-690).  Do_intersect_3(Bbox_3, Triangle_3)
// (Bbox_3_Triangle_3_do_intersect.h: separating axes) is replaced by a conservative test — the binary32 box of the query, rounded outward, must overlap the
// child's box — and the leaves decide exactly (geom.cuh::tri_tri_do_intersect), so results do not depend on the tree.  Same while-while scheme as traverse_line.
// SEG: the primitives are Segment_3 (stored as the records (source, target, target), b200aabb_set_segments): the leaf test is
// Do_intersect_3(Triangle_3 query, Segment_3 datum) (Segment_3_Triangle_3_do_intersect.h:177-262)
template <int MODE, class XF = NoXf, bool SEG = false>
B2_HD uint32_t traverse_triangle(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root, const Tri3& query, TravStat& st,
                                 uint32_t* list = nullptr, uint32_t list_cap = 0, const XF& xf = XF()) {
    constexpr bool kStopAtFirst = MODE == kAnyHit || MODE == kAnyId;
    const BoxCtx c = make_box_of(query);
    uint32_t found = 0, nlisted = 0, found_id = kNone;
    auto leaf = [&](uint32_t ref) -> bool {
        uint32_t id;
        const Tri3 t = xf.apply(load_tri(tris, ref & ~kLeafBit, id));
        ++st.leaves;
        if (!(SEG ? seg_tri_do_intersect(query, t.a, t.b, st.nexact) : tri_tri_do_intersect(query, t, st.nexact))) return false;
        if (MODE == kList) { if (nlisted < list_cap) list[nlisted] = id; ++nlisted; }
        if (MODE == kAnyId) found_id = id;
        return true;
    };
    if (root & kLeafBit) found = leaf(root) ? 1u : 0u;
    else {
        uint32_t stack[kStackCap];
        constexpr int kPendCount = 4, kPendGoAt = kPendCount - 1;
        uint32_t pend[kPendCount + 1];
        int sp = 0, np = 0;
        uint32_t cur = root;
        bool done = false;
        while (!done) {
            while (cur != kNone && np < kPendGoAt) {          // node phase
                const NodeView n = load_node(nodes, cur);
                ++st.nodes;
                const bool h0 = xf.box_meets(n.lo0, n.hi0, c), h1 = xf.box_meets(n.lo1, n.hi1, c);
                uint32_t next = kNone;
                if (h0) { if (n.c0 & kLeafBit) pend[np++] = n.c0; else next = n.c0; }
                if (h1) {
                    if (n.c1 & kLeafBit) pend[np++] = n.c1;
                    else if (next == kNone) next = n.c1;
                    else if (sp < kStackCap) stack[sp++] = n.c1;
                    else st.overflow = 1;
                }
                if (next == kNone && sp > 0) next = stack[--sp];
                cur = next;
            }
            while (np > 0) {                                  // leaf phase
                if (leaf(pend[--np])) { ++found; if (kStopAtFirst) { done = true; break; } }
            }
            if (cur == kNone) break;
        }
    }
    return MODE == kAnyId ? found_id : (MODE == kAnyHit ? (found ? 1u : 0u) : found);
}

// Side_of_triangle_mesh (Polygon_mesh_processing/include/CGAL/Side_of_triangle_mesh.h:265-306, internal/Side_of_triangle_mesh/
// Point_inside_vertical_ray_cast.h:69-121, Ray_3_Triangle_3_traversal_traits.h:30-76): one ray from the point; every primitive it crosses in the
// interior of a facet counts, the source lying on a primitive ends the query (ON_BOUNDARY), and a ray that passes through an edge or a vertex or
// lies in a primitive's plane makes the attempt INDETERMINATE (the caller takes another ray).  Returns the number of crossings, kSideBoundary or
// kSideIndeterminate.
constexpr int kSideBoundary = -1, kSideIndeterminate = -2;
// XF: the mesh stands under an affine transformation (does_A_contains_a_CC_of_B, Rigid_triangle_mesh_collision_detection.h:131-147, with the
// Transformed_tree_helper of AABB_traversal_traits_with_transformation.h:190-224): the primitives' images are classified, node boxes are mapped conservatively
template <class XF = NoXf>
B2_HD int traverse_side(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root, const Vec3& s, const Vec3& q, TravStat& st, const XF& xf = XF()) {
    const LineCtx c = make_line(s, q);
    int crossings = 0, status = 0;
    traverse_line<kRayQ, true>(nodes, root, c, st, [&](uint32_t ref) -> bool {
        uint32_t id;
        const Tri3 t = xf.apply(load_tri(tris, ref & ~kLeafBit, id));
        ++st.leaves;
        const int k = ray_tri_classify(t, s, q, st.nexact);
        if (k < 0) return false;
        if (k == 0) { ++crossings; return false; }
        status = k == 4 ? kSideBoundary : kSideIndeterminate;
        return true;                                         // m_stop
    }, xf);
    return status ? status : crossings;
}

// Plane_3 queries (a x + b y + c z + d = 0): Do_intersect_3(Plane_3, Bbox_3) (Iso_cuboid_3_Plane_3_do_intersect.h:90-125: the signs of the two
// corners extreme along the normal) is replaced by a conservative binary64 test — the plane can only meet the box if its value at the box
// centre is within the box's half extent along the normal, plus the rounding of that evaluation — and the leaves decide exactly.
B2_HD bool plane_box(const float* lo, const float* hi, double a, double b, double c, double d) {
    const double cx = 0.5 * ((double)lo[0] + (double)hi[0]), cy = 0.5 * ((double)lo[1] + (double)hi[1]), cz = 0.5 * ((double)lo[2] + (double)hi[2]);
    const double hx = 0.5 * ((double)hi[0] - (double)lo[0]), hy = 0.5 * ((double)hi[1] - (double)lo[1]), hz = 0.5 * ((double)hi[2] - (double)lo[2]);
    const double s = ((a * cx + b * cy) + c * cz) + d;
    const double r = (fabs(a) * hx + fabs(b) * hy) + fabs(c) * hz;
    const double mag = ((fabs(a) * (fabs(cx) + hx) + fabs(b) * (fabs(cy) + hy)) + fabs(c) * (fabs(cz) + hz)) + fabs(d);
    return !(fabs(s) > r + 1e-14 * mag);      // (NaN / overflow: keep the box)
}
template <bool kStopAtFirst, class Leaf>
B2_HD uint32_t traverse_plane_nodes(const BvhNode* __restrict__ nodes, uint32_t root, double a, double b, double c, double d, TravStat& st, Leaf&& leaf) {
    uint32_t found = 0;
    if (root & kLeafBit) found = leaf(root) ? 1u : 0u;
    else {
        uint32_t stack[kStackCap];
        int sp = 0;
        uint32_t cur = root;
        while (cur != kNone) {
            const NodeView n = load_node(nodes, cur);
            ++st.nodes;
            const bool h0 = plane_box(n.lo0, n.hi0, a, b, c, d), h1 = plane_box(n.lo1, n.hi1, a, b, c, d);
            uint32_t next = kNone;
            if (h0) { if (n.c0 & kLeafBit) { if (leaf(n.c0)) { ++found; if (kStopAtFirst) break; } } else next = n.c0; }
            if (h1) {
                if (n.c1 & kLeafBit) { if (leaf(n.c1)) { ++found; if (kStopAtFirst) break; } }
                else if (next == kNone) next = n.c1;
                else if (sp < kStackCap) stack[sp++] = n.c1;
                else st.overflow = 1;
            }
            if (next == kNone && sp > 0) next = stack[--sp];
            cur = next;
        }
    }
    return found;
}
template <int MODE>
B2_HD uint32_t traverse_plane(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root, double a, double b, double c, double d,
                              TravStat& st, uint32_t* list = nullptr, uint32_t list_cap = 0) {
    uint32_t nlisted = 0, found_id = kNone;
    const uint32_t found = traverse_plane_nodes<MODE == kAnyHit || MODE == kAnyId>(nodes, root, a, b, c, d, st, [&](uint32_t ref) -> bool {
        uint32_t id;
        const Tri3 t = load_tri(tris, ref & ~kLeafBit, id);
        ++st.leaves;
        if (!plane_tri_do_intersect(t, a, b, c, d, st.nexact)) return false;
        if (MODE == kList) { if (nlisted < list_cap) list[nlisted] = id; ++nlisted; }
        if (MODE == kAnyId) found_id = id;
        return true;
    });
    return MODE == kAnyId ? found_id : (MODE == kAnyHit ? (found ? 1u : 0u) : found);
}
// the intersection OBJECTS of a Plane_3 query (Intersect_3(Plane_3, Triangle_3)): records as traverse_ray_objects; kind 3 = the primitive's triangle itself
template <int MODE>
B2_HD uint32_t traverse_plane_objects(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root, double a, double b, double c, double d,
                                      TravStat& st, uint32_t* ids, uint8_t* kinds, double* pts, uint32_t cap) {
    uint32_t nlisted = 0;
    return traverse_plane_nodes<MODE == kAnyId>(nodes, root, a, b, c, d, st, [&](uint32_t ref) -> bool {
        uint32_t id;
        const Tri3 t = load_tri(tris, ref & ~kLeafBit, id);
        ++st.leaves;
        const SegX x = plane_tri_intersection(t, a, b, c, d, st.nexact);
        if (x.kind == 0) return false;
        if (MODE != kCount) {
            if (nlisted < cap) {
                ids[nlisted] = id; kinds[nlisted] = (uint8_t)x.kind;
                double* o = pts + 6 * (uint64_t)nlisted;
                const bool pt = x.kind == 1 || x.kind == 2, seg = x.kind == 2;
                o[0] = pt ? x.p0.x : 0.0; o[1] = pt ? x.p0.y : 0.0; o[2] = pt ? x.p0.z : 0.0;
                o[3] = seg ? x.p1.x : 0.0; o[4] = seg ? x.p1.y : 0.0; o[5] = seg ? x.p1.z : 0.0;
            }
            ++nlisted;
        }
        return true;
    });
}

// all_intersections / any_intersection (Ray_3): every primitive whose Intersect_3(Ray_3, Triangle_3) is non-empty, with the
// intersection object (Listing_intersection_traits / First_intersection_traits, AABB_traversal_traits.h:58-144; AABB_tree.h:958-985).
// The object is the one first_intersection constructs (geom.cuh::ray_tri_intersection: the reference evaluates the construction-side
// predicates on the ROUNDED point a + (q - p), so this set can differ from do_intersect's for a ray nearly parallel to a triangle,
// SURVEY.md §7 hard part 3).  Entry k of the query: ids[k], kinds[k] (1 Point_3, 2 Segment_3), pts[6k..] = point / two endpoints.
// MODE kCount: count only; kList: fill (at most cap entries); kAnyId: stop at the first, written as entry 0.
// KIND kSegmentQ: the same for Segment_3 queries with Intersect_3(Segment_3, Triangle_3) (geom.cuh::seg_tri_intersection).
template <int MODE, int KIND = kRayQ>
B2_HD uint32_t traverse_ray_objects(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root, const Vec3& s, const Vec3& q,
                                    TravStat& st, uint32_t* ids, uint8_t* kinds, double* pts, uint32_t cap) {
    // Line_3 (s = point, q = direction vector): the reference's Intersect_3(Line_3, Triangle_3) decides with the ROUNDED second point s + q (it is a
    // construction), while its node test is the exact predicate on (s, q): within rounding of a primitive's border the reference's own answer therefore
    // depends on its tree.  Here nodes and primitives see the same line, through s and fl(s + q): the result is every primitive whose Intersect_3 is
    // non-empty, whatever the tree (the reference's result is a subset of it, equal unless the line grazes a primitive within rounding).
    const LineCtx c = KIND == kLineQ ? make_line(s, s + q) : make_line(s, q);
    uint32_t nlisted = 0;
    return traverse_line<KIND, MODE == kAnyId>(nodes, root, c, st, [&](uint32_t ref) -> bool {
        uint32_t id;
        const Tri3 t = load_tri(tris, ref & ~kLeafBit, id);
        ++st.leaves;
        struct { int kind; Vec3 p0, p1; } x;
        if (KIND == kSegmentQ || KIND == kLineQ) {
            const SegX y = KIND == kSegmentQ ? seg_tri_intersection(t, s, q, st.nexact) : line_tri_intersection(t, s, q, st.nexact);   // (Line_3: q is the direction vector)
            x.kind = y.kind; x.p0 = y.p0; x.p1 = y.p1;
        } else {
            int nex = 0;
            const RayX y = ray_tri_intersection_call(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, s.x, s.y, s.z, q.x, q.y, q.z, c.major, &nex);
            st.nexact += nex;
            x.kind = y.kind; x.p0 = y.p0; x.p1 = y.p1;
        }
        if (x.kind == 0) return false;
        if (MODE != kCount) {
            if (nlisted < cap) {
                ids[nlisted] = id; kinds[nlisted] = (uint8_t)x.kind;
                double* o = pts + 6 * (uint64_t)nlisted;
                o[0] = x.p0.x; o[1] = x.p0.y; o[2] = x.p0.z;
                const bool seg = x.kind == 2;
                o[3] = seg ? x.p1.x : 0.0; o[4] = seg ? x.p1.y : 0.0; o[5] = seg ? x.p1.z : 0.0;
            }
            ++nlisted;
        }
        return true;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// first_intersection(Ray_3)   (AABB_ray_intersection.h:49-225)
// Result = the intersected primitive with the smallest ray parameter t as the reference computes it.  On exactly
// equal t the smallest primitive id wins and `tie` is set (the reference keeps whichever it popped first).
// `near` is set when another intersected primitive's t lies within the ROUNDING UNCERTAINTY of the winner's (RayX::u;
// large for a ray grazing a triangle): the reference compares the same rounded values and prunes boxes with the inexact
// slab distance (AABB_traits_3.h:89-125, AABB_ray_intersection.h:78), so in that band its own answer depends on the
// order in which its tree presents the candidates.  Boxes are pruned with a 2e-6 relative slack so that such
// candidates are seen.  The stack keeps each deferred node's conservative entry parameter: stale entries cost no load.
// ---------------------------------------------------------------------------------------------------------------
struct FirstHit { uint32_t prim; int kind; double t; Vec3 p0, p1; int tie; int near; };

// skip_id / skip_mask: first_intersection(query, skip) (AABB_tree.h:307-328): primitives the caller's functor rejects are ignored — here the
// one primitive skip_id (kNone: none) and those whose bit is set in skip_mask (bit per primitive id, mask_words 32-bit words; nullptr: none)
template <bool SKIP = false>
B2_HD FirstHit traverse_first(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root,
                              const Vec3& s, const Vec3& q, double mag, TravStat& st, uint32_t skip_id = kNone,
                              const uint32_t* __restrict__ skip_mask = nullptr, uint32_t mask_words = 0) {
    const LineCtx c = make_line(s, q, mag);
    FirstHit best; best.prim = kNone; best.kind = 0; best.t = DBL_MAX; best.p0 = s; best.p1 = s; best.tie = 0; best.near = 0;
    double best_u = 0.0;
    double strict_lo = DBL_MAX;   // smallest (t - u) among the intersected primitives whose t is LARGER than the winner's
    double tie_lo = DBL_MAX;      // smallest (t - u) among those whose t EQUALS the winner's
    bool tie = false;
    float tlimit = INFINITY;      // prune boxes whose entry parameter certainly exceeds the best hit (rounded up)
    auto leaf = [&](uint32_t ref) {
        uint32_t id;
        const Tri3 t = load_tri(tris, ref & ~kLeafBit, id);
        ++st.leaves;
        // skip(primitive.id()): AABB_ray_intersection.h:86,99,113 — NOT consulted for a one-primitive tree, which the reference answers directly (:214)
        if (SKIP && !(root & kLeafBit) && (id == skip_id || (skip_mask && (id >> 5) < mask_words && ((skip_mask[id >> 5] >> (id & 31u)) & 1u)))) return;
        int nex = 0;
        const RayX x = ray_tri_intersection_call(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, s.x, s.y, s.z, q.x, q.y, q.z, c.major, &nex);
        st.nexact += nex;
        if (x.kind == 0) return;
        if (x.t < best.t || (x.t == best.t && id < best.prim)) {
            if (best.prim != kNone) {
                if (x.t < best.t) { strict_lo = fmin(strict_lo, fmin(best.t - best_u, tie_lo)); tie = false; tie_lo = DBL_MAX; }   // the old winner and its ties fall behind
                else { tie = true; tie_lo = fmin(tie_lo, best.t - best_u); }                                                       // same t, smaller id
            }
            best.prim = id; best.kind = x.kind; best.t = x.t; best.p0 = x.p0; best.p1 = x.p1; best_u = x.u;
            tlimit = float_above(x.t + ((fabs(x.t) * 2e-6 + x.u) + c.t_abs_slack));
        } else if (x.t == best.t) { tie = true; tie_lo = fmin(tie_lo, x.t - x.u); }
        else strict_lo = fmin(strict_lo, x.t - x.u);
    };
    if (root & kLeafBit) { leaf(root); }
    else {
        uint32_t stack[kStackCap];
        float stack_t[kStackCap];     // lower bound of the deferred node's entry parameter
        int sp = 0;
        uint32_t pend[2];             // leaves awaiting the exact test, nearer box first
        float pend_t[2];
        int np = 0;
        uint32_t cur = root;
        float cur_t = 0.0f;
        while (true) {
            while (cur != kNone && np == 0) {             // node phase
                const NodeView n = load_node(nodes, cur);
                ++st.nodes;
                float t0, t1;
                bool h0 = line_box(n.lo0, n.hi0, c, tlimit, t0);
                bool h1 = line_box(n.lo1, n.hi1, c, tlimit, t1);
                uint32_t c0 = n.c0, c1 = n.c1;
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
            for (int k = 0; k < np; ++k)                   // leaf phase
                if (pend_t[k] <= tlimit) leaf(pend[k]);
            np = 0;
            if (cur != kNone && cur_t > tlimit) {          // the node chosen before the leaf phase may have fallen behind the new hit
                cur = kNone;
                while (cur == kNone && sp > 0) {
                    --sp;
                    if (stack_t[sp] <= tlimit) { cur = stack[sp]; cur_t = stack_t[sp]; }
                }
            }
            if (cur == kNone) break;
        }
    }
    if (best.prim != kNone) {
        best.tie = tie;                                    // another primitive attains exactly the winner's t
        best.near = strict_lo <= best.t + best_u;          // another one, with a larger rounded t, lies within the rounding uncertainty (both may hold)
    }
    return best;
}

// ---------------------------------------------------------------------------------------------------------------
// closest_point_and_primitive(Point_3)   (AABB_tree.h:987-1055, Projection_traits AABB_traversal_traits.h:277-327)
// Result = the candidate (per-triangle projected point, rounded exactly as the reference rounds it) with the smallest
// EXACT squared distance to the query; on exactly equal distance the smallest primitive id wins and `tie` is set;
// (`near` is never set by this query, see the end of the function.)
// A box is skipped only when it is certainly farther than the incumbent by the margin of `widen` (1e-13 relative plus a
// few ulps of the coordinate magnitude, since candidates are ROUNDED points), which is wider than the reference's own
// semi-static margin (AABB_traits_3.h:318-459): every candidate the reference could have preferred is examined.
//
// Three phases per round: (A) nodes, recording leaf children whose box reaches into the incumbent ball; (B1) a rigorous
// division-free binary32 pre-test of the recorded leaves (tangent plane of the incumbent ball, direction of the triangle's
// centroid); (B2) for the survivors the sharper pre-test and then the binary64 projection of the reference + the exact
// comparison.  B2 is postponed until a thread has collected kSurvGo survivors (or has no incumbent yet, or is finishing) and
// then runs for every thread of the warp that has a survivor.
// ---------------------------------------------------------------------------------------------------------------
struct Closest { Vec3 p; uint32_t prim; double sqd; int tie; int near; };

constexpr int kPendGo = 3;    // leave the node phase once this many leaves are pending (a node adds at most two)
constexpr int kSurvGo = 2;    // run the binary64 phase once this many pre-test survivors are waiting

B2_HD double widen_sqdist(double sqd, double delta) { return sqd + ((2.0 * delta * sqrt(sqd) + delta * delta) + sqd * 1e-13); }

// (explicit FMAs: the library is compiled with -fmad=false for the binary64 constructions; the binary32 bounds only gain accuracy)
B2_HD float dot3f(const float* u, const float* v) { return fmaf(u[2], v[2], fmaf(u[1], v[1], u[0] * v[0])); }

// Approximate closest point of the triangle (a,b,c) to the ORIGIN in binary32 (region walk of Ericson, "Real-Time
// Collision Detection" 5.1.5, specialised to p = 0).  Only used as a direction for the rigorous bound below, so its own
// rounding errors (even gross ones on needle triangles) can only weaken that bound, never invalidate it.
B2_HD void approx_closest_to_origin(const float* a, const float* b, const float* c, float* u) {
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float d1 = -dot3f(ab, a), d2 = -dot3f(ac, a);
    const float d3 = -dot3f(ab, b), d4 = -dot3f(ac, b);
    const float d5 = -dot3f(ab, c), d6 = -dot3f(ac, c);
    float v = 0.0f, w = 0.0f;                              // result = a + v ab + w ac
    const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    if (d1 <= 0.0f && d2 <= 0.0f) { }                                               // vertex a
    else if (d3 >= 0.0f && d4 <= d3) { v = 1.0f; }                                  // vertex b
    else if (d6 >= 0.0f && d5 <= d6) { w = 1.0f; }                                  // vertex c
    else if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) { v = d1 / (d1 - d3); }        // edge ab
    else if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) { w = d2 / (d2 - d6); }        // edge ac
    else if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) { w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); v = 1.0f - w; }   // edge bc
    else { const float den = 1.0f / (va + vb + vc); v = vb * den; w = vc * den; }   // interior
    u[0] = a[0] + (v * ab[0] + w * ac[0]); u[1] = a[1] + (v * ab[1] + w * ac[1]); u[2] = a[2] + (v * ab[2] + w * ac[2]);
}

// RIGOROUS lower-bound test: is the triangle certainly farther from the query than `rbound` (>= sqrt of the pruning bound)?
// ra/rb/rc = triangle vertices relative to the query, rounded to binary32 (each coordinate off by <= m * 2^-23.9,
// m = largest magnitude among them).  For ANY direction u every point x of the triangle satisfies
//   |x| >= (u.x)/|u| >= min_i (u.v_i)/|u|     (x is a convex combination of the vertices),
// so the triangle is rejected when min_i fl(u.r_i) - E > rbound * |u| with E = 5e-7 m |u| covering the coordinate
// rounding and the three binary32 dot products; `su` = fl(|u|) (2e-7 relative) and the final comparison carry 1e-6.
B2_HD bool certainly_farther(const float* u, float su, const float* ra, const float* rb, const float* rc, float m, float rbound) {
    const float md = fminf(dot3f(u, ra), fminf(dot3f(u, rb), dot3f(u, rc)));
    // (range guard as closest_warp.cuh::wp_lower: no binary32 overflow of the dot products, and they and E stay above the denormal range)
    return su > 1e-18f && su < 1e19f && m < 1e18f && m * su > 1e-30f && md - 5e-7f * m * su > rbound * su * 1.000001f;
}

// binary32 pre-test of one primitive on coordinates RELATIVE to the query o (the subtraction is done in binary64): true = the
// primitive is certainly farther than rbound (>= sqrt of the widened incumbent distance).  ub / sub = direction query ->
// incumbent closest point and its length.  sharp = also try the direction of the triangle's own approximate closest point.
B2_HD bool tri_certainly_outside(const Tri3& t, const Vec3& o, const float* ub, float sub, float rbound, bool sharp) {
    const float ra[3] = {(float)(t.a.x - o.x), (float)(t.a.y - o.y), (float)(t.a.z - o.z)};
    const float rb[3] = {(float)(t.b.x - o.x), (float)(t.b.y - o.y), (float)(t.b.z - o.z)};
    const float rc[3] = {(float)(t.c.x - o.x), (float)(t.c.y - o.y), (float)(t.c.z - o.z)};
    const float m = fmaxf(fmaxf(fmaxf(fabsf(ra[0]), fabsf(ra[1])), fmaxf(fabsf(ra[2]), fabsf(rb[0]))),
                          fmaxf(fmaxf(fabsf(rb[1]), fabsf(rb[2])), fmaxf(fmaxf(fabsf(rc[0]), fabsf(rc[1])), fabsf(rc[2]))));
    // (1) tangent plane of the incumbent ball: all three vertices beyond it
    if (certainly_farther(ub, sub, ra, rb, rc, m, rbound)) return true;
    // (2) direction of the triangle's centroid: nearly optimal for a triangle that is small compared with its distance
    const float g[3] = {(ra[0] + rb[0]) + rc[0], (ra[1] + rb[1]) + rc[1], (ra[2] + rb[2]) + rc[2]};
    if (certainly_farther(g, sqrtf(dot3f(g, g)), ra, rb, rc, m, rbound)) return true;
    if (!sharp) return false;
    // (3) direction of this triangle's own (approximate) closest point
    float u[3];
    approx_closest_to_origin(ra, rb, rc, u);
    return certainly_farther(u, sqrtf(dot3f(u, u)), ra, rb, rc, m, rbound);
}

// SEG: Segment_3 primitives stored as (source, target, target): the binary32 pre-tests bound the distance to that flat triangle = the segment, the candidate is
// the reference's projection on the SEGMENT
template <bool SEG = false>
B2_HD Closest traverse_closest(const BvhNode* __restrict__ nodes, const TriRec* __restrict__ tris, uint32_t root,
                               const Vec3& o, double mag, TravStat& st) {
    Closest best; best.p = o; best.prim = kNone; best.sqd = DBL_MAX; best.tie = 0; best.near = 0;
    const double delta = 4e-16 * fmax(mag, fmax3(fabs(o.x), fabs(o.y), fabs(o.z)));
    const PointCtx pc = make_point(o);
    float bound = INFINITY;       // prune boxes certainly farther than the incumbent (widened, rounded up)
    float rbound = INFINITY;      // >= sqrt(bound)
    float ub[3] = {0.0f, 0.0f, 0.0f}, sub = 0.0f;   // direction query -> incumbent closest point and its length (binary32)
    auto certainly_outside = [&](const Tri3& t, bool sharp) -> bool { return tri_certainly_outside(t, o, ub, sub, rbound, sharp); };
    // the reference's per-primitive step: Construct_projected_point_3 + Compare_distance_3 against the incumbent
    auto evaluate = [&](const Tri3& t, uint32_t id) {
        const Vec3 cp = SEG ? closest_candidate_segment(t.a, t.b, o, st.nexact) : closest_candidate(t, o, st.nexact);
        const Vec3 v = cp - o;
        const double sqd = (v.x * v.x + v.y * v.y) + v.z * v.z;   // Compute_squared_distance_3(query, closest): AABB_tree.h:1024-1032
        bool take;
        if (best.prim == kNone) take = true;
        else {
            const int sg = cmp_dist(o, cp, best.p, st.nexact);     // Compare_distance_3(query, candidate, incumbent): AABB_traits_3.h:724-732
            if (sg == 0) { take = id < best.prim; best.tie = 1; }
            else { take = sg < 0; if (take) best.tie = 0; }
        }
        if (take) {
            best.p = cp; best.prim = id; best.sqd = sqd;
            const double w = widen_sqdist(sqd, delta);
            bound = float_above(w);
            rbound = float_above(sqrt(w));
            ub[0] = (float)v.x; ub[1] = (float)v.y; ub[2] = (float)v.z;
            sub = sqrtf(dot3f(ub, ub));
        }
    };
    if (root & kLeafBit) {
        uint32_t id;
        const Tri3 t = load_tri(tris, root & ~kLeafBit, id);
        ++st.leaves;
        evaluate(t, id);
        return best;
    }
    uint32_t stack[kStackCap];
    float stack_d[kStackCap];     // lower bound of the deferred node's squared distance
    int sp = 0;
    uint32_t pend[kPendGo + 1];   // leaves whose box reaches into the ball, awaiting the binary32 pre-test
    int np = 0;
    uint32_t surv[kSurvGo + kPendGo + 1];   // pre-test survivors awaiting the binary64 projection
    int ns = 0;
    uint32_t cur = root;
    float cur_d = 0.0f;
    while (true) {
        // (A) node phase.  Without an incumbent the first pending leaf ends it: its projection gives the first ball.
        const int go = best.prim == kNone ? 1 : kPendGo;
        while (cur != kNone && np < go) {
            const NodeView n = load_node(nodes, cur);
            ++st.nodes;
            B2_TRACE_NODE(cur);
            float d0 = point_box_sqdist(n.lo0, n.hi0, pc), d1 = point_box_sqdist(n.lo1, n.hi1, pc);
            uint32_t c0 = n.c0, c1 = n.c1;
            if (d1 < d0) { const uint32_t tc = c0; c0 = c1; c1 = tc; const float td = d0; d0 = d1; d1 = td; }
            uint32_t next = kNone;
            if (d0 <= bound) { if (c0 & kLeafBit) pend[np++] = c0; else { next = c0; cur_d = d0; } }
            if (d1 <= bound) {
                if (c1 & kLeafBit) pend[np++] = c1;
                else if (next == kNone) { next = c1; cur_d = d1; }
                else if (sp < kStackCap) { stack[sp] = c1; stack_d[sp] = d1; ++sp; }
                else st.overflow = 1;
            }
            while (next == kNone && sp > 0) {
                --sp;
                if (stack_d[sp] <= bound) { next = stack[sp]; cur_d = stack_d[sp]; }
            }
            cur = next;
        }
        // (B1) cheap rigorous pre-test of the pending leaves
        for (int k = 0; k < np; ++k) {
            uint32_t id;
            const Tri3 t = load_tri(tris, pend[k] & ~kLeafBit, id);
            ++st.leaves;
            if (best.prim != kNone && certainly_outside(t, false)) { ++st.culled; continue; }
            surv[ns++] = pend[k];
        }
        np = 0;
        // (B2) binary64 projection of the survivors
        const bool due = ns > 0 && (best.prim == kNone || ns >= kSurvGo || cur == kNone);
        if (warp_any(due)) {
            while (ns > 0) {
                --ns;
                uint32_t id;
                const Tri3 t = load_tri(tris, surv[ns] & ~kLeafBit, id);
                if (best.prim != kNone && certainly_outside(t, true)) { ++st.culled; continue; }
                evaluate(t, id);
            }
            if (cur != kNone && cur_d > bound) {           // the node chosen before this phase may have fallen outside the new ball
                cur = kNone;
                while (cur == kNone && sp > 0) {
                    --sp;
                    if (stack_d[sp] <= bound) { cur = stack[sp]; cur_d = stack_d[sp]; }
                }
            }
        }
        if (cur == kNone && ns == 0) break;
    }
    // no near-tie band here: the reference never prunes when unsure and compares candidates exactly, so its answer is
    // order-independent except on exact ties (measured: 0 other mismatches in 10^6 queries; candidates on a shared edge agree
    // to a few ulps for a third of all queries and are decided identically by both)
    return best;
}

}  // namespace b2
