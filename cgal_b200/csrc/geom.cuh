// geom.cuh — per-primitive device routines of the traversal kernels: triangle tests (exact) and the plain-double
// constructions whose operation order must match the reference bit for bit.
// Compile with -fmad=false: no FMA contraction may happen in the constructions.
#pragma once
#include <float.h>
#include "predicates.cuh"

namespace b2 {

struct Tri3 { Vec3 a, b, c; };

B2_HD Vec3 cross3(const Vec3& v, const Vec3& w) {   // Construct_cross_product_vector_3, Cartesian/function_objects.h:2263-2268
    return {v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x};
}
B2_HD double dot3(const Vec3& v, const Vec3& w) { return (v.x * w.x + v.y * w.y) + v.z * w.z; }
B2_HD double sqlen3(const Vec3& v) { return (v.x * v.x + v.y * v.y) + v.z * v.z; }

// supporting plane a x + b y + c z + d = 0 of (P,Q,R): plane_from_pointsC3, constructions/kernel_ftC3.h:225-241
struct Plane4 { double a, b, c, d; };
B2_HD Plane4 plane_of(const Tri3& t) {
    const Vec3 rp = t.a - t.c, rq = t.b - t.c;
    Plane4 h;
    h.a = rp.y * rq.z - rq.y * rp.z;
    h.b = rp.z * rq.x - rq.z * rp.x;
    h.c = rp.x * rq.y - rq.x * rp.y;
    h.d = ((-h.a * t.c.x) - h.b * t.c.y) - h.c * t.c.z;
    return h;
}

// "the line (x,y) passes the triangle on the non-positive side of all three edges"
// = the three Orientation_3 tests shared by all straddling cases of Segment/Ray x Triangle
B2_HD bool passes_edges(const Vec3& x, const Vec3& y, const Tri3& t, int& nexact) {
    return orient3(x, y, t.a, t.b, nexact) != 1 && orient3(x, y, t.b, t.c, nexact) != 1 && orient3(x, y, t.c, t.a, nexact) != 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Coplanar configurations (rare).  The 27-way (pqa,pqb,pqc) case analyses of the reference are stored as tables
// indexed by (pqa+1)*9 + (pqb+1)*3 + (pqc+1); vertices are numbered 0=a 1=b 2=c AFTER the ccw swap.
// ---------------------------------------------------------------------------------------------------------------
// Segment x Triangle (Intersections_3/internal/Segment_3_Triangle_3_do_intersect.h:24-152): each entry is two edge
// tests "coplanar_orientation(edge, endpoint) != NEGATIVE"; edge e: 0=(a,b) 1=(b,c) 2=(c,a); endpoint 0=p 1=q.
// code = e1 | pt1<<2 | e2<<3 | pt2<<5 ; 0xFF = no intersection
#define B2_ST(e1, p1, e2, p2) ((e1) | ((p1) << 2) | ((e2) << 3) | ((p2) << 5))
__constant__ static const unsigned char kSegTriCoplanar[27] = {
    /* pqa=-1 */ /* pqb=-1 */ 0xFF, B2_ST(1, 0, 2, 1), B2_ST(1, 0, 2, 1),
                 /* pqb= 0 */ B2_ST(0, 0, 1, 1), B2_ST(0, 0, 2, 1), B2_ST(0, 0, 2, 1),
                 /* pqb=+1 */ B2_ST(0, 0, 1, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 2, 1),
    /* pqa= 0 */ /* pqb=-1 */ B2_ST(0, 1, 2, 0), B2_ST(0, 1, 1, 0), B2_ST(0, 1, 1, 0),
                 /* pqb= 0 */ B2_ST(1, 1, 2, 0), B2_ST(1, 1, 2, 0), B2_ST(1, 0, 2, 1),
                 /* pqb=+1 */ B2_ST(0, 0, 1, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 2, 1),
    /* pqa=+1 */ /* pqb=-1 */ B2_ST(0, 1, 2, 0), B2_ST(0, 1, 2, 0), B2_ST(0, 1, 1, 0),
                 /* pqb= 0 */ B2_ST(0, 1, 2, 0), B2_ST(0, 1, 2, 0), B2_ST(0, 1, 1, 0),
                 /* pqb=+1 */ B2_ST(1, 1, 2, 0), B2_ST(1, 1, 2, 0), 0xFF};
// Ray x Triangle do_intersect (Ray_3_Triangle_3_do_intersect.h:75-191): one test "coplanar_orientation(x,y,p) != POSITIVE";
// code: 0 = (a,c)  1 = (c,b)  2 = (b,a) ; 0xFF = no intersection
__constant__ static const unsigned char kRayTriCoplanar[27] = {
    /* pqa=-1 */ 0xFF, 1, 1, /**/ 2, 2, 2, /**/ 2, 2, 2,
    /* pqa= 0 */ 0, 1, 1, /**/ 0, 0, 1, /**/ 2, 2, 2,
    /* pqa=+1 */ 0, 0, 1, /**/ 0, 0, 1, /**/ 0, 0, 0xFF};
// Ray x Triangle intersection (Ray_3_Triangle_3_intersection.h:147-396):
//   kind 0 none | 1 clip(x,y,z,negative_side) | 2 vertex v | 3 aligned edge (u -> w)
//   code = kind | arg<<2 ; clip arg = rot | neg<<2 with rot 0=(a,b,c) 1=(b,c,a) 2=(c,a,b); vertex arg = v; edge arg = u | w<<2
#define B2_CLIP(rot, neg) (1 | (((rot) | ((neg) << 2)) << 2))
#define B2_VERT(v) (2 | ((v) << 2))
#define B2_EDGE(u, w) (3 | (((u) | ((w) << 2)) << 2))
__constant__ static const unsigned char kRayTriIntersectCoplanar[27] = {
    /* pqa=-1 */ /* pqb=-1 */ 0, B2_VERT(2), B2_CLIP(0, 0),
                 /* pqb= 0 */ B2_VERT(1), B2_EDGE(2, 1), B2_CLIP(1, 1),
                 /* pqb=+1 */ B2_CLIP(2, 0), B2_CLIP(2, 0), B2_CLIP(1, 1),
    /* pqa= 0 */ /* pqb=-1 */ B2_VERT(0), B2_EDGE(0, 2), B2_CLIP(2, 1),
                 /* pqb= 0 */ B2_EDGE(1, 0), 0, B2_EDGE(0, 1),
                 /* pqb=+1 */ B2_CLIP(2, 0), B2_EDGE(2, 0), B2_VERT(0),
    /* pqa=+1 */ /* pqb=-1 */ B2_CLIP(1, 0), B2_CLIP(1, 0), B2_CLIP(2, 1),
                 /* pqb= 0 */ B2_CLIP(1, 0), B2_EDGE(1, 2), B2_VERT(1),
                 /* pqb=+1 */ B2_CLIP(0, 1), B2_VERT(2), 0};

#if !defined(__CUDA_ARCH__)
// host-side copies for the CPU self-check of the device headers (tests/hostcheck); same initialisers
namespace host_tables {
static const unsigned char SegTri[27] = {
    0xFF, B2_ST(1, 0, 2, 1), B2_ST(1, 0, 2, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 2, 1), B2_ST(0, 0, 2, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 2, 1),
    B2_ST(0, 1, 2, 0), B2_ST(0, 1, 1, 0), B2_ST(0, 1, 1, 0), B2_ST(1, 1, 2, 0), B2_ST(1, 1, 2, 0), B2_ST(1, 0, 2, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 1, 1), B2_ST(0, 0, 2, 1),
    B2_ST(0, 1, 2, 0), B2_ST(0, 1, 2, 0), B2_ST(0, 1, 1, 0), B2_ST(0, 1, 2, 0), B2_ST(0, 1, 2, 0), B2_ST(0, 1, 1, 0), B2_ST(1, 1, 2, 0), B2_ST(1, 1, 2, 0), 0xFF};
static const unsigned char RayTri[27] = {0xFF, 1, 1, 2, 2, 2, 2, 2, 2, 0, 1, 1, 0, 0, 1, 2, 2, 2, 0, 0, 1, 0, 0, 1, 0, 0, 0xFF};
static const unsigned char RayTriX[27] = {
    0, B2_VERT(2), B2_CLIP(0, 0), B2_VERT(1), B2_EDGE(2, 1), B2_CLIP(1, 1), B2_CLIP(2, 0), B2_CLIP(2, 0), B2_CLIP(1, 1),
    B2_VERT(0), B2_EDGE(0, 2), B2_CLIP(2, 1), B2_EDGE(1, 0), 0, B2_EDGE(0, 1), B2_CLIP(2, 0), B2_EDGE(2, 0), B2_VERT(0),
    B2_CLIP(1, 0), B2_CLIP(1, 0), B2_CLIP(2, 1), B2_CLIP(1, 0), B2_EDGE(1, 2), B2_VERT(1), B2_CLIP(0, 1), B2_VERT(2), 0};
}  // namespace host_tables
#endif

B2_HD unsigned char tbl_segtri(int i) {
#if defined(__CUDA_ARCH__)
    return kSegTriCoplanar[i];
#else
    return host_tables::SegTri[i];
#endif
}
B2_HD unsigned char tbl_raytri(int i) {
#if defined(__CUDA_ARCH__)
    return kRayTriCoplanar[i];
#else
    return host_tables::RayTri[i];
#endif
}
B2_HD unsigned char tbl_raytrix(int i) {
#if defined(__CUDA_ARCH__)
    return kRayTriIntersectCoplanar[i];
#else
    return host_tables::RayTriX[i];
#endif
}

struct CoplanarFrame { Vec3 v[3]; int idx; };   // ccw-ordered vertices + table index
B2_HD CoplanarFrame coplanar_frame(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    CoplanarFrame f;
    const bool ccw = coplanar_orient(t.a, t.b, t.c, nexact) == 1;
    f.v[0] = t.a; f.v[1] = ccw ? t.b : t.c; f.v[2] = ccw ? t.c : t.b;
    const int pqa = coplanar_orient(p, q, f.v[0], nexact), pqb = coplanar_orient(p, q, f.v[1], nexact), pqc = coplanar_orient(p, q, f.v[2], nexact);
    f.idx = (pqa + 1) * 9 + (pqb + 1) * 3 + (pqc + 1);
    return f;
}

B2_HD_NOINLINE bool seg_tri_coplanar(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    const CoplanarFrame f = coplanar_frame(t, p, q, nexact);
    const unsigned code = tbl_segtri(f.idx);
    if (code == 0xFF) return false;
    const int e1 = code & 3, e2 = (code >> 3) & 3;
    const Vec3& x1 = ((code >> 2) & 1) ? q : p;
    const Vec3& x2 = ((code >> 5) & 1) ? q : p;
    return coplanar_orient(f.v[e1], f.v[(e1 + 1) % 3], x1, nexact) != -1 && coplanar_orient(f.v[e2], f.v[(e2 + 1) % 3], x2, nexact) != -1;
}
B2_HD_NOINLINE bool ray_tri_coplanar(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    const CoplanarFrame f = coplanar_frame(t, p, q, nexact);
    const unsigned code = tbl_raytri(f.idx);
    if (code == 0xFF) return false;
    const int x = code == 0 ? 0 : (code == 1 ? 2 : 1), y = code == 0 ? 2 : (code == 1 ? 1 : 0);
    return coplanar_orient(f.v[x], f.v[y], p, nexact) != 1;
}

// Do_intersect_3(Segment_3, Triangle_3): Segment_3_Triangle_3_do_intersect.h:177-262
B2_HD bool seg_tri_do_intersect(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    const int sp = orient3(t.a, t.b, t.c, p, nexact);
    const int sq = orient3(t.a, t.b, t.c, q, nexact);
    if (sp != 0 ? sq == sp : sq == 0) {                 // same open side, or both on the plane
        if (sp != 0) return false;
        return seg_tri_coplanar(t, p, q, nexact);
    }
    const bool fwd = sp > 0 || (sp == 0 && sq < 0);     // which endpoint sees the triangle counterclockwise
    return fwd ? passes_edges(p, q, t, nexact) : passes_edges(q, p, t, nexact);
}

// Do_intersect_3(Ray_3, Triangle_3): Ray_3_Triangle_3_do_intersect.h:193-317 (exact ray direction)
B2_HD bool ray_tri_do_intersect(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    const int dir = orient3_dir(t.a, t.b, t.c, p, q, nexact);
    const int sp = orient3(t.a, t.b, t.c, p, nexact);
    if (dir == 0) return sp == 0 ? ray_tri_coplanar(t, p, q, nexact) : false;
    if (sp == dir) return false;                        // source on one side, heading away
    const bool fwd = sp > 0 || (sp == 0 && dir < 0);
    return fwd ? passes_edges(p, q, t, nexact) : passes_edges(q, p, t, nexact);
}

// The same with r3t3_do_intersect_endpoint_position_visitor (Ray_3_Triangle_3_do_intersect.h:23-62): HOW the ray meets the triangle.
// -1 no intersection; 0 CROSS_FACET; 1 CROSS_SEGMENT; 2 CROSS_VERTEX (the ray passes through an edge / a vertex: one / two of the three edge
// orientations vanish); 3 COPLANAR_RAY; 4 ENDPOINT_IN_TRIANGLE (the source lies in the triangle's plane, on the closed triangle).
B2_HD int ray_tri_classify(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    const int dir = orient3_dir(t.a, t.b, t.c, p, q, nexact);
    const int sp = orient3(t.a, t.b, t.c, p, nexact);
    if (dir == 0) return sp == 0 && ray_tri_coplanar(t, p, q, nexact) ? 3 : -1;
    if (sp == dir) return -1;
    const bool fwd = sp > 0 || (sp == 0 && dir < 0);
    const Vec3& x = fwd ? p : q; const Vec3& y = fwd ? q : p;
    int zeros = 0;
    const int o0 = orient3(x, y, t.a, t.b, nexact);
    if (o0 == 1) return -1;
    zeros += o0 == 0;
    const int o1 = orient3(x, y, t.b, t.c, nexact);
    if (o1 == 1) return -1;
    zeros += o1 == 0;
    const int o2 = orient3(x, y, t.c, t.a, nexact);
    if (o2 == 1) return -1;
    zeros += o2 == 0;
    return sp == 0 ? 4 : zeros;
}

// Do_intersect_3(Line_3, Triangle_3): Line_3_Triangle_3_do_intersect.h:27-79.  The line is (point p, direction vector v) as CGAL stores it;
// "q" of the reference is p + v, formed exactly (the predicate runs in Filtered_predicate's exact kernel when the filter fails).
B2_HD bool line_tri_do_intersect(const Tri3& t, const Vec3& p, const Vec3& v, int& nexact) {
    // orientation(a,b,c,p) != COPLANAR || orientation(a,b,c,q) != COPLANAR; with the first zero, the second is the sign of det[b-a | c-a | v]
    const Vec3 zero = {0.0, 0.0, 0.0};
    if (orient3(t.a, t.b, t.c, p, nexact) != 0 || orient3_dir(t.a, t.b, t.c, zero, v, nexact) != 0) {
        const int pqab = orient3_vec(v, p, t.a, t.b, nexact), pqbc = orient3_vec(v, p, t.b, t.c, nexact);
        if (pqab > 0) return pqbc != -1 && orient3_vec(v, p, t.c, t.a, nexact) != -1;
        if (pqab < 0) return pqbc != 1 && orient3_vec(v, p, t.c, t.a, nexact) != 1;
        if (pqbc > 0) return orient3_vec(v, p, t.c, t.a, nexact) != -1;
        if (pqbc < 0) return orient3_vec(v, p, t.c, t.a, nexact) != 1;
        return true;
    }
    const int pqa = coplanar_orient_vec(p, v, t.a, nexact);                 // the line lies in the triangle's plane (:76-77)
    return coplanar_orient_vec(p, v, t.b, nexact) != pqa || coplanar_orient_vec(p, v, t.c, nexact) != pqa;
}
// Do_intersect_3(Plane_3, Triangle_3): Plane_3_Triangle_3_do_intersect.h:27-51; the plane is its four coefficients (a, b, c, d)
B2_HD bool plane_tri_do_intersect(const Tri3& t, double a, double b, double c, double d, int& nexact) {
    const int s0 = plane_side(a, b, c, d, t.a, nexact);
    if (s0 == 0) return true;
    return plane_side(a, b, c, d, t.b, nexact) != s0 || plane_side(a, b, c, d, t.c, nexact) != s0;
}

// ---------------------------------------------------------------------------------------------------------------
// Do_intersect_3(Triangle_3, Triangle_3): Triangle_3_Triangle_3_do_intersect.h:154-450 (Guigue & Devillers).  The reference's nested switches over the
// signs of one triangle's vertices against the other's plane select, for each triangle, the two edges that cross the other plane ("min" and "max" edge);
// here the 27 sign combinations index one table, the same for both triangles.  Entry = s_min | t_min << 2 | s_max << 4 | t_max << 6 (vertex numbers of
// the triangle), 0xFF = all vertices strictly on one side (no intersection), 0xFE = all three on the plane (coplanar triangles).
// ---------------------------------------------------------------------------------------------------------------
#define B2_TT(smin, tmin, smax, tmax) ((smin) | ((tmin) << 2) | ((smax) << 4) | ((tmax) << 6))
#define B2_TT_TABLE { \
    /* d0=-1 */ 0xFF, B2_TT(2, 0, 1, 2), B2_TT(2, 0, 1, 2), B2_TT(1, 2, 0, 1), B2_TT(2, 0, 0, 1), B2_TT(2, 0, 0, 1), B2_TT(1, 2, 0, 1), B2_TT(2, 0, 0, 1), B2_TT(2, 0, 0, 1), \
    /* d0= 0 */ B2_TT(0, 1, 2, 0), B2_TT(0, 1, 1, 2), B2_TT(0, 1, 1, 2), B2_TT(1, 2, 2, 0), 0xFE, B2_TT(2, 0, 1, 2), B2_TT(1, 2, 0, 1), B2_TT(1, 2, 0, 1), B2_TT(2, 0, 0, 1), \
    /* d0=+1 */ B2_TT(0, 1, 2, 0), B2_TT(0, 1, 2, 0), B2_TT(0, 1, 1, 2), B2_TT(0, 1, 2, 0), B2_TT(0, 1, 2, 0), B2_TT(0, 1, 1, 2), B2_TT(1, 2, 2, 0), B2_TT(1, 2, 2, 0), 0xFF}
__constant__ static const unsigned char kTriTriEdges[27] = B2_TT_TABLE;
#if !defined(__CUDA_ARCH__)
namespace host_tables { static const unsigned char TriTri[27] = B2_TT_TABLE; }
#endif
B2_HD unsigned char tbl_tritri(int i) {
#if defined(__CUDA_ARCH__)
    return kTriTriEdges[i];
#else
    return host_tables::TriTri[i];
#endif
}
B2_HD const Vec3& tri_vertex(const Tri3& t, int i) { return i == 0 ? t.a : (i == 1 ? t.b : t.c); }

// Coplanar triangles (do_intersect_coplanar, :100-152, decides by locating one vertex among the other triangle's edge lines; this is the separating-axis
// form of the same exact decision): two closed, non-flat triangles of one plane are disjoint exactly when the supporting line of an edge of one of them
// has the other triangle strictly on its outer side.  Both are first made counterclockwise as the reference does (:120-130).  Flat triangles are outside
// the reference's precondition (:108-109) and unspecified here too.
B2_HD_NOINLINE bool tri_tri_coplanar(const Tri3& t1, const Tri3& t2, int& nexact) {
    Vec3 u[3] = {t1.a, t1.b, t1.c}, w[3] = {t2.a, t2.b, t2.c};
    if (coplanar_orient(u[0], u[1], u[2], nexact) == -1) { const Vec3 x = u[1]; u[1] = u[2]; u[2] = x; }
    if (coplanar_orient(w[0], w[1], w[2], nexact) == -1) { const Vec3 x = w[1]; w[1] = w[2]; w[2] = x; }
    for (int e = 0; e < 3; ++e) {
        const Vec3 &a = u[e], &b = u[(e + 1) % 3];
        if (coplanar_orient(a, b, w[0], nexact) == -1 && coplanar_orient(a, b, w[1], nexact) == -1 && coplanar_orient(a, b, w[2], nexact) == -1) return false;
    }
    for (int e = 0; e < 3; ++e) {
        const Vec3 &a = w[e], &b = w[(e + 1) % 3];
        if (coplanar_orient(a, b, u[0], nexact) == -1 && coplanar_orient(a, b, u[1], nexact) == -1 && coplanar_orient(a, b, u[2], nexact) == -1) return false;
    }
    return true;
}

// t1 = the query (p, q, r), t2 = the primitive (a, b, c): the order of the reference's call do_intersect(query, datum)
B2_HD_NOINLINE bool tri_tri_do_intersect(const Tri3& t1, const Tri3& t2, int& nexact) {
    const int dp = orient3(t2.a, t2.b, t2.c, t1.a, nexact), dq = orient3(t2.a, t2.b, t2.c, t1.b, nexact), dr = orient3(t2.a, t2.b, t2.c, t1.c, nexact);   // :176-178
    const unsigned e1 = tbl_tritri((dp + 1) * 9 + (dq + 1) * 3 + (dr + 1));
    if (e1 == 0xFF) return false;
    if (e1 == 0xFE) return tri_tri_coplanar(t1, t2, nexact);
    const int da = orient3(t1.a, t1.b, t1.c, t2.a, nexact), db = orient3(t1.a, t1.b, t1.c, t2.b, nexact), dc = orient3(t1.a, t1.b, t1.c, t2.c, nexact);   // :312-314
    const unsigned e2 = tbl_tritri((da + 1) * 9 + (db + 1) * 3 + (dc + 1));
    if (e2 == 0xFF) return false;
    if (e2 == 0xFE) return tri_tri_coplanar(t1, t2, nexact);       // (only reachable with a flat t1, :424-428)
    // the two triangles cut the common line of their planes in two intervals; they overlap iff (:446-447)
    return orient3(tri_vertex(t1, e1 & 3), tri_vertex(t1, (e1 >> 2) & 3), tri_vertex(t2, e2 & 3), tri_vertex(t2, (e2 >> 2) & 3), nexact) != 1 &&
           orient3(tri_vertex(t1, (e1 >> 4) & 3), tri_vertex(t1, (e1 >> 6) & 3), tri_vertex(t2, (e2 >> 6) & 3), tri_vertex(t2, (e2 >> 4) & 3), nexact) != 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Intersect_3(Ray_3, Triangle_3) + the ray parameter of AABB_ray_intersection.h:168-198
// ---------------------------------------------------------------------------------------------------------------
struct RayX { int kind; Vec3 p0, p1; double t; double u; };   // kind 0 none, 1 point, 2 segment; u = rounding uncertainty of t

B2_HD double ray_param(const Vec3& s, const Vec3& d, int ax, const Vec3& x) { return (comp(x, ax) - comp(s, ax)) / comp(d, ax); }
B2_HD int ray_major_axis(const Vec3& d) {
    int m = 0;
    if (fabs(d.y) > fabs(d.x)) m = 1;
    if (fabs(d.z) > fabs(comp(d, m))) m = 2;
    return m;
}

// t3r3_intersection_coplanar_aux (point version), Ray_3_Triangle_3_intersection.h:28-61
B2_HD Vec3 coplanar_lines_meet(const Vec3& p, const Vec3& v, const Vec3& a, const Vec3& b) {
    const Vec3 ab = b - a, pa = a - p;
    const Vec3 pa_ab = cross3(pa, ab), v_ab = cross3(v, ab);
    const double t = dot3(pa_ab, v_ab) / sqlen3(v_ab);
    return {p.x + v.x * t, p.y + v.y * t, p.z + v.z * t};
}

B2_HD_NOINLINE RayX ray_tri_intersection_coplanar(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    RayX r; r.kind = 0; r.t = 0; r.u = 0; r.p0 = p; r.p1 = p;
    const CoplanarFrame f = coplanar_frame(t, p, q, nexact);
    const unsigned code = tbl_raytrix(f.idx);
    const int kind = code & 3, arg = code >> 2;
    if (kind == 0) return r;
    if (kind == 2) {                                    // the line through p,q touches the triangle at one vertex
        const Vec3& x = f.v[arg];
        if (collinear_ordered(p, x, q) || collinear_ordered(p, q, x)) { r.kind = 1; r.p0 = x; }
        return r;
    }
    if (kind == 3) {                                    // the ray runs along edge u -> w
        const Vec3& u = f.v[arg & 3]; const Vec3& w = f.v[arg >> 2];
        r.kind = 2; r.p0 = collinear_ordered(p, u, w) ? u : p; r.p1 = w;
        return r;
    }
    // clip the ray into the triangle: :63-146, vertices rotated so that `c` is the isolated one
    const int rot = arg & 3; const bool neg = (arg >> 2) != 0;
    const Vec3& a = f.v[rot]; const Vec3& b = f.v[(rot + 1) % 3]; const Vec3& c = f.v[(rot + 2) % 3];
    const int cap = neg ? coplanar_orient(c, a, p, nexact) : coplanar_orient(b, c, p, nexact);
    if (cap < 0) return r;
    if (cap == 0) { r.kind = 1; r.p0 = p; return r; }
    const Vec3 v = q - p;
    r.kind = 2; r.p0 = p;
    if (neg) {
        if (coplanar_orient(b, c, p, nexact) == -1) r.p0 = coplanar_lines_meet(p, v, b, c);
        r.p1 = coplanar_lines_meet(p, v, c, a);
    } else {
        if (coplanar_orient(c, a, p, nexact) == -1) r.p0 = coplanar_lines_meet(p, v, c, a);
        r.p1 = coplanar_lines_meet(p, v, b, c);
    }
    return r;
}

// Ray_3_Triangle_3_intersection.h:417-539; the hit point is Line_3 x Plane_3 (Line_3_Plane_3_intersection.h:52-85)
B2_HD RayX ray_tri_intersection(const Tri3& t, const Vec3& s, const Vec3& q, const Vec3& d, int major_axis, int& nexact) {
    RayX r; r.kind = 0; r.t = 0; r.u = 0; r.p0 = s; r.p1 = s;
    const Vec3 shifted = t.a + d;                         // Ray_3(a, q - s).second_point(): a ROUNDED point (:439)
    const int dir = orient3(t.a, t.b, t.c, shifted, nexact);
    const int sp = orient3(t.a, t.b, t.c, s, nexact);
    if (sp == 0 && dir == 0) {
        r = ray_tri_intersection_coplanar(t, s, q, nexact);
        if (r.kind == 1) r.t = ray_param(s, d, major_axis, r.p0);
        else if (r.kind == 2) r.t = fmin(ray_param(s, d, major_axis, r.p0), ray_param(s, d, major_axis, r.p1));
        r.u = 1.8e-15 * fabs(r.t);
        return r;
    }
    if (dir == 0 || sp == dir) return r;                  // parallel off-plane, or heading away
    const bool fwd = sp > 0 || (sp == 0 && dir < 0);
    if (!(fwd ? passes_edges(s, q, t, nexact) : passes_edges(q, s, t, nexact))) return r;
    const Plane4 h = plane_of(t);
    const double num = ((h.a * s.x + h.b * s.y) + h.c * s.z) + h.d;
    const double den = (h.a * d.x + h.b * d.y) + h.c * d.z;
    if (den == 0.0) return r;                             // the reference then produces no Point_3
    Vec3 x = {den * s.x - num * d.x, den * s.y - num * d.y, den * s.z - num * d.z};
    if (den != 1.0) { x.x = x.x / den; x.y = x.y / den; x.z = x.z / den; }   // Cartesian/Vector_3.h:73-75
    r.kind = 1; r.p0 = x; r.p1 = x; r.t = ray_param(s, d, major_axis, x);
    // t = -num/den in exact arithmetic: its rounding error grows like |plane terms| / |den| for a grazing ray.  The
    // reference compares these rounded values, so inside this band its first hit depends on its traversal order.
    r.u = 1.8e-15 * (2.0 * (((fabs(h.a * s.x) + fabs(h.b * s.y)) + fabs(h.c * s.z)) + fabs(h.d)) / fabs(den) + fabs(r.t));
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Intersect_3(Segment_3, Triangle_3): Intersections_3/internal/Segment_3_Triangle_3_intersection.h:28-569 (kind 0 none, 1 Point_3 p0, 2 Segment_3 p0 p1)
// ---------------------------------------------------------------------------------------------------------------
struct SegX { int kind; Vec3 p0, p1; };
// :63-133  clip [p,q] into the triangle; c is isolated on the negative side of (p,q)
B2_HD SegX t3s3_clip(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& p, const Vec3& q, bool negative_side, int& nexact) {
    SegX r; r.kind = 0; r.p0 = p; r.p1 = p;
    const int bcq = coplanar_orient(b, c, q, nexact), cap = coplanar_orient(c, a, p, nexact);
    if (bcq == -1 || cap == -1) return r;
    if (bcq == 0) { r.kind = 1; r.p0 = q; return r; }
    if (cap == 0) { r.kind = 1; r.p0 = p; return r; }
    Vec3 pe = p, qe = q;
    const Vec3 v = q - p;
    if (coplanar_orient(b, c, p, nexact) == -1) pe = coplanar_lines_meet(p, v, b, c);     // t3s3_intersection_coplanar_aux (point version), :28-61
    if (coplanar_orient(c, a, q, nexact) == -1) qe = coplanar_lines_meet(p, v, c, a);
    r.kind = 2; r.p0 = negative_side ? pe : qe; r.p1 = negative_side ? qe : pe;
    return r;
}
// :135-186  [a,b] and [p,q] collinear, same direction
B2_HD SegX t3s3_collinear(const Vec3& a, const Vec3& b, const Vec3& p, const Vec3& q) {
    SegX r; r.kind = 0; r.p0 = p; r.p1 = p;
    if (collinear_ordered(p, a, b)) {
        if (collinear_ordered(a, b, q)) { r.kind = 2; r.p0 = a; r.p1 = b; }
        else if (collinear_ordered(q, a, b)) { if (same_point(a, q)) { r.kind = 1; r.p0 = a; } }
        else { r.kind = 2; r.p0 = a; r.p1 = q; }
        return r;
    }
    if (collinear_ordered(p, b, q)) {
        if (same_point(p, b)) { r.kind = 1; r.p0 = p; } else { r.kind = 2; r.p0 = p; r.p1 = b; }
        return r;
    }
    if (collinear_ordered(a, b, p)) { if (same_point(p, b)) { r.kind = 1; r.p0 = p; } return r; }
    r.kind = 2; r.p0 = p; r.p1 = q;
    return r;
}
// :188-398  the 27 sign patterns of (pqa, pqb, pqc)
B2_HD_NOINLINE SegX seg_tri_intersection_coplanar(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    SegX none; none.kind = 0; none.p0 = p; none.p1 = p;
    const bool ccw = coplanar_orient(t.a, t.b, t.c, nexact) == 1;
    const Vec3 a = t.a, b = ccw ? t.b : t.c, c = ccw ? t.c : t.b;
    const int pqa = coplanar_orient(p, q, a, nexact), pqb = coplanar_orient(p, q, b, nexact), pqc = coplanar_orient(p, q, c, nexact);
    const int code = (pqa + 1) * 9 + (pqb + 1) * 3 + (pqc + 1);
    const Vec3* vtx = nullptr;                      // "vertex x if it lies inside [p,q]"
    switch (code) {
        case 26: case 0: case 13: return none;                                  // (+,+,+), (-,-,-), degenerate (0,0,0)
        case 24: return t3s3_clip(a, b, c, p, q, true, nexact);                 // (+,+,-)
        case 25: case 1: vtx = &c; break;                                       // (+,+,0), (-,-,0)
        case 20: return t3s3_clip(c, a, b, p, q, true, nexact);                 // (+,-,+)
        case 18: case 19: return t3s3_clip(b, c, a, q, p, false, nexact);       // (+,-,-), (+,-,0)
        case 23: case 3: vtx = &b; break;                                       // (+,0,+), (-,0,-)
        case 21: return t3s3_clip(b, c, a, q, p, false, nexact);                // (+,0,-)
        case 22: return t3s3_collinear(b, c, p, q);                             // (+,0,0)
        case 8: return t3s3_clip(b, c, a, p, q, true, nexact);                  // (-,+,+)
        case 6: case 7: return t3s3_clip(c, a, b, q, p, false, nexact);         // (-,+,-), (-,+,0)
        case 2: return t3s3_clip(a, b, c, q, p, false, nexact);                 // (-,-,+)
        case 5: return t3s3_clip(b, c, a, p, q, true, nexact);                  // (-,0,+)
        case 4: return t3s3_collinear(c, b, p, q);                              // (-,0,0)
        case 17: case 9: vtx = &a; break;                                       // (0,+,+), (0,-,-)
        case 15: return t3s3_clip(c, a, b, q, p, false, nexact);                // (0,+,-)
        case 16: return t3s3_collinear(c, a, p, q);                             // (0,+,0)
        case 11: return t3s3_clip(c, a, b, p, q, true, nexact);                 // (0,-,+)
        case 10: return t3s3_collinear(a, c, p, q);                             // (0,-,0)
        case 14: return t3s3_collinear(a, b, p, q);                             // (0,0,+)
        case 12: return t3s3_collinear(b, a, p, q);                             // (0,0,-)
        default: return none;
    }
    if (collinear_ordered(p, *vtx, q)) { none.kind = 1; none.p0 = *vtx; }
    return none;
}
// :400-548
B2_HD SegX seg_tri_intersection(const Tri3& t, const Vec3& p, const Vec3& q, int& nexact) {
    SegX r; r.kind = 0; r.p0 = p; r.p1 = p;
    const int sp = orient3(t.a, t.b, t.c, p, nexact), sq = orient3(t.a, t.b, t.c, q, nexact);
    if (sp == 0 && sq == 0) return seg_tri_intersection_coplanar(t, p, q, nexact);
    if (sp == sq) return r;
    // which endpoint sees the triangle counterclockwise: p when it is on the positive side, or in the plane with q on the negative side
    const bool fwd = sp > 0 || (sp == 0 && sq < 0);
    if (!(fwd ? passes_edges(p, q, t, nexact) : passes_edges(q, p, t, nexact))) return r;
    if (sp == 0) { r.kind = 1; r.p0 = p; return r; }
    if (sq == 0) { r.kind = 1; r.p0 = q; return r; }
    const Plane4 h = plane_of(t);                          // Line_3 x Plane_3, Line_3_Plane_3_intersection.h:52-85, as for the ray
    const Vec3 d = q - p;
    const double num = ((h.a * p.x + h.b * p.y) + h.c * p.z) + h.d;
    const double den = (h.a * d.x + h.b * d.y) + h.c * d.z;
    if (den == 0.0) return r;
    Vec3 x = {den * p.x - num * d.x, den * p.y - num * d.y, den * p.z - num * d.z};
    if (den != 1.0) { x.x = x.x / den; x.y = x.y / den; x.z = x.z / den; }
    r.kind = 1; r.p0 = x;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Intersect_3(Line_3, Triangle_3): Line_3_Triangle_3_intersection.h:27-419, and Intersect_3(Plane_3, Triangle_3): Plane_3_Triangle_3_intersection.h:27-146.
// Constructions: plain doubles with filtered predicates on the points they form — point_on(l, 1) is the ROUNDED p + v here, unlike in the predicate
// line_tri_do_intersect above.  Object kinds as SegX, plus 3 = the Triangle_3 itself (Plane_3 queries, all three vertices on the plane; no points).
// ---------------------------------------------------------------------------------------------------------------
// Intersect_3(Plane_3 (a, b, c, d), Line_3 (point s, direction d)): Line_3_Plane_3_intersection.h:52-85; false: no Point_3
B2_HD bool plane_line_point(double pa, double pb, double pc, double pd, const Vec3& s, const Vec3& d, Vec3& out) {
    const double num = ((pa * s.x + pb * s.y) + pc * s.z) + pd;
    const double den = (pa * d.x + pb * d.y) + pc * d.z;
    if (den == 0.0) return false;
    Vec3 x = {den * s.x - num * d.x, den * s.y - num * d.y, den * s.z - num * d.z};
    if (den != 1.0) { x.x = x.x / den; x.y = x.y / den; x.z = x.z / den; }
    out = x;
    return true;
}
B2_HD SegX t3l3_clip(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& p, const Vec3& v, bool negative_side) {      // :66-88
    const Vec3 l_bc = coplanar_lines_meet(p, v, b, c), l_ca = coplanar_lines_meet(p, v, c, a);
    SegX r; r.kind = 2; r.p0 = negative_side ? l_bc : l_ca; r.p1 = negative_side ? l_ca : l_bc;
    return r;
}
B2_HD_NOINLINE SegX line_tri_intersection_coplanar(const Tri3& t, const Vec3& p, const Vec3& q, const Vec3& v, int& nexact) {   // :90-277
    SegX none; none.kind = 0; none.p0 = p; none.p1 = p;
    const bool ccw = coplanar_orient(t.a, t.b, t.c, nexact) == 1;
    const Vec3 a = t.a, b = ccw ? t.b : t.c, c = ccw ? t.c : t.b;
    const int pqa = coplanar_orient(p, q, a, nexact), pqb = coplanar_orient(p, q, b, nexact), pqc = coplanar_orient(p, q, c, nexact);
    SegX r = none;
    switch ((pqa + 1) * 9 + (pqb + 1) * 3 + (pqc + 1)) {
        case 26: case 0: case 13: return none;                                   // (+,+,+), (-,-,-), degenerate
        case 24: return t3l3_clip(a, b, c, p, v, true);                          // (+,+,-)
        case 25: case 1: r.kind = 1; r.p0 = c; return r;                         // (+,+,0), (-,-,0)
        case 20: return t3l3_clip(c, a, b, p, v, true);                          // (+,-,+)
        case 18: case 19: case 21: return t3l3_clip(b, c, a, p, v, false);       // (+,-,-), (+,-,0), (+,0,-)
        case 23: case 3: r.kind = 1; r.p0 = b; return r;                         // (+,0,+), (-,0,-)
        case 22: r.kind = 2; r.p0 = b; r.p1 = c; return r;                       // (+,0,0)
        case 8: case 5: return t3l3_clip(b, c, a, p, v, true);                   // (-,+,+), (-,0,+)
        case 6: case 7: case 15: return t3l3_clip(c, a, b, p, v, false);         // (-,+,-), (-,+,0), (0,+,-)
        case 2: return t3l3_clip(a, b, c, p, v, false);                          // (-,-,+)
        case 4: r.kind = 2; r.p0 = c; r.p1 = b; return r;                        // (-,0,0)
        case 17: case 9: r.kind = 1; r.p0 = a; return r;                         // (0,+,+), (0,-,-)
        case 16: r.kind = 2; r.p0 = c; r.p1 = a; return r;                       // (0,+,0)
        case 11: return t3l3_clip(c, a, b, p, v, true);                          // (0,-,+)
        case 10: r.kind = 2; r.p0 = a; r.p1 = c; return r;                       // (0,-,0)
        case 14: r.kind = 2; r.p0 = a; r.p1 = b; return r;                       // (0,0,+)
        case 12: r.kind = 2; r.p0 = b; r.p1 = a; return r;                       // (0,0,-)
        default: return none;
    }
}
B2_HD SegX line_tri_intersection(const Tri3& t, const Vec3& p, const Vec3& v, int& nexact) {                                  // :279-400
    SegX r; r.kind = 0; r.p0 = p; r.p1 = p;
    const Vec3 q = p + v;                                   // point_on(l, 1): rounded
    if (orient3(t.a, t.b, t.c, p, nexact) == 0 && orient3(t.a, t.b, t.c, q, nexact) == 0) return line_tri_intersection_coplanar(t, p, q, v, nexact);
    const int pqab = orient3(p, q, t.a, t.b, nexact), pqbc = orient3(p, q, t.b, t.c, nexact);
    bool through;
    if (pqab > 0) through = pqbc != -1 && orient3(p, q, t.c, t.a, nexact) != -1;
    else if (pqab < 0) through = pqbc != 1 && orient3(p, q, t.c, t.a, nexact) != 1;
    else if (pqbc > 0) through = orient3(p, q, t.c, t.a, nexact) != -1;
    else if (pqbc < 0) through = orient3(p, q, t.c, t.a, nexact) != 1;
    else through = true;
    if (!through) return r;
    const Plane4 h = plane_of(t);
    Vec3 x;
    if (plane_line_point(h.a, h.b, h.c, h.d, p, v, x)) { r.kind = 1; r.p0 = x; }
    return r;
}
B2_HD SegX plane_tri_intersection(const Tri3& t, double pa, double pb, double pc, double pd, int& nexact) {
    SegX r; r.kind = 0; r.p0 = t.a; r.p1 = t.a;
    const int o0 = plane_side(pa, pb, pc, pd, t.a, nexact), o1 = plane_side(pa, pb, pc, pd, t.b, nexact), o2 = plane_side(pa, pb, pc, pd, t.c, nexact);
    Vec3 x;
    if (o0 == 0) {
        if (o1 == 0) { if (o2 == 0) { r.kind = 3; return r; } r.kind = 2; r.p0 = t.a; r.p1 = t.b; return r; }
        if (o2 == 0) { r.kind = 2; r.p0 = t.a; r.p1 = t.c; return r; }
        if (o1 == o2) { r.kind = 1; r.p0 = t.a; return r; }
        if (plane_line_point(pa, pb, pc, pd, t.b, t.c - t.b, x)) { r.kind = 2; r.p0 = x; r.p1 = t.a; }
        return r;
    }
    if (o1 == 0) {
        if (o2 == 0) { r.kind = 2; r.p0 = t.b; r.p1 = t.c; return r; }
        if (o2 == o0) { r.kind = 1; r.p0 = t.b; return r; }
        if (plane_line_point(pa, pb, pc, pd, t.a, t.c - t.a, x)) { r.kind = 2; r.p0 = x; r.p1 = t.b; }
        return r;
    }
    if (o2 == 0) {
        if (o1 == o0) { r.kind = 1; r.p0 = t.c; return r; }
        if (plane_line_point(pa, pb, pc, pd, t.a, t.b - t.a, x)) { r.kind = 2; r.p0 = x; r.p1 = t.c; }
        return r;
    }
    Vec3 first = t.a, last = t.a;
    int np = 0;
    if (o0 != o1) { if (!plane_line_point(pa, pb, pc, pd, t.a, t.b - t.a, x)) return r; first = x; last = x; ++np; }
    if (o0 != o2) { if (!plane_line_point(pa, pb, pc, pd, t.a, t.c - t.a, x)) return r; if (np == 0) first = x; last = x; ++np; }
    if (o1 != o2) { if (!plane_line_point(pa, pb, pc, pd, t.b, t.c - t.b, x)) return r; if (np == 0) first = x; last = x; ++np; }
    if (np == 0) return r;
    r.kind = 2; r.p0 = first; r.p1 = last;
    return r;
}

// The same routine OUT OF LINE with its arguments in registers, for the traversal kernels: a ray meets ~1.3 leaves, and keeping this long
// routine (five orientation predicates, a division chain) out of the traversal loop lets the loop run at 8 blocks per SM without spilling.
B2_HD_NOINLINE RayX ray_tri_intersection_call(double ax, double ay, double az, double bx, double by, double bz, double cx, double cy, double cz,
                                              double sx, double sy, double sz, double qx, double qy, double qz, int major_axis, int* nexact_out) {
    int nexact = 0;
    const Vec3 s = {sx, sy, sz}, q = {qx, qy, qz};
    const RayX r = ray_tri_intersection(Tri3{{ax, ay, az}, {bx, by, bz}, {cx, cy, cz}}, s, q, q - s, major_axis, nexact);
    *nexact_out = nexact;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// Construct_projected_point_3(Triangle_3, Point_3): Kernel_23/include/CGAL/Kernel/function_objects.h:3193-3263
// ---------------------------------------------------------------------------------------------------------------
B2_HD Vec3 foot_on_line(const Vec3& lp, const Vec3& ld, const Vec3& p) {   // Cartesian/function_objects.h:3200-3217
    const Vec3 dp = p - lp;
    const double lambda = ((ld.x * dp.x + ld.y * dp.y) + ld.z * dp.z) / ((ld.x * ld.x + ld.y * ld.y) + ld.z * ld.z);
    return {lp.x + lambda * ld.x, lp.y + lambda * ld.y, lp.z + lambda * ld.z};
}
B2_HD_NOINLINE Vec3 closest_on_segment(const Vec3& s, const Vec3& e, const Vec3& query) {   // function_objects.h:3266-3285
    if (same_point(s, e)) return s;
    const Vec3 v = e - s;
    if (dot3(v, query - s) <= 0) return s;
    if (dot3(v, query - e) >= 0) return e;
    return foot_on_line(s, v, query);
}
B2_HD_NOINLINE Vec3 closest_on_degenerate_triangle(const Tri3& t, const Vec3& o) {          // :3208-3247
    const Vec3 ab = t.b - t.a, ac = t.c - t.a, bc = t.c - t.b;
    const double lab = fmax(fmax(ab.x, ab.y), ab.z), lac = fmax(fmax(ac.x, ac.y), ac.z), lbc = fmax(fmax(bc.x, bc.y), bc.z);
    if (lab > lac) return lab > lbc ? closest_on_segment(t.a, t.b, o) : closest_on_segment(t.b, t.c, o);
    return lac > lbc ? closest_on_segment(t.a, t.c, o) : closest_on_segment(t.b, t.c, o);
}
B2_HD Vec3 closest_on_triangle(const Tri3& t, const Vec3& o) {
    const Plane4 h = plane_of(t);
    if (h.a == 0.0 && h.b == 0.0 && h.c == 0.0) return closest_on_degenerate_triangle(t, o);
    const double num = ((h.a * o.x + h.b * o.y) + h.c * o.z) + h.d;       // projection_planeC3, kernel_ftC3.h:296-312
    const double den = (h.a * h.a + h.b * h.b) + h.c * h.c;
    const double lambda = num / den;
    const Vec3 foot = {o.x - lambda * h.a, o.y - lambda * h.b, o.z - lambda * h.c};
    const Vec3 w = {h.a, h.b, h.c};
    bool outside = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {                                         // edges (a,b) (b,c) (c,a): :3130-3132
        const Vec3& p1 = i == 0 ? t.a : (i == 1 ? t.b : t.c);
        const Vec3& p2 = i == 0 ? t.b : (i == 1 ? t.c : t.a);
        const Vec3 e = p2 - p1, u = foot - p1;
        if (dot3(cross3(e, u), w) < 0) {                                  // foot is outside this edge (:3051-3052)
            if (dot3(u, e) >= 0 && dot3(foot - p2, p1 - p2) >= 0) return foot_on_line(p1, e, foot);
            outside = true;
        }
    }
    if (!outside) return foot;
    const double d1 = sqlen3(t.a - foot), d2 = sqlen3(t.b - foot), d3 = sqlen3(t.c - foot);   // nearest_point_3 :3075-3101
    if (d2 >= d1 && d3 >= d1) return t.a;
    return d3 >= d2 ? t.b : t.c;
}

// The candidate a primitive contributes to closest_point: its projection, or its reference point (first vertex) when that is
// strictly nearer.  The reference starts every query from best_hint() = the nearest reference point over all primitives
// (AABB_tree.h:591-612, internal/AABB_search_tree.h) and keeps it unless a projection is not farther (AABB_traits_3.h:724-732);
// the nearest reference point belongs to a primitive whose box reaches the final ball, so taking the reference point of every
// visited primitive into account reproduces the hint's effect exactly.  It only ever matters for degenerate (zero-area)
// triangles, where Construct_projected_point_3 may return a point that is not the closest one (function_objects.h:3208-3247).
// (arguments by value: a call through references would route the nine coordinates through local memory)
B2_HD_NOINLINE int cmp_dist_out_of_line(double px, double py, double pz, double qx, double qy, double qz, double rx, double ry, double rz) {
    int nexact = 0;
    const int s = cmp_dist(Vec3{px, py, pz}, Vec3{qx, qy, qz}, Vec3{rx, ry, rz}, nexact);
    return (s + 1) | (nexact ? 4 : 0);
}
B2_HD Vec3 closest_candidate(const Tri3& t, const Vec3& o, int& nexact) {
    const Vec3 cp = closest_on_triangle(t, o);
    // plain binary64 first: both squared distances are sums of squares (relative error < 1e-15), so a reference point that looks
    // farther by more than 1e-12 relative IS farther; only the rest (the closest point is at or next to the first vertex, or a
    // degenerate triangle) takes the exact comparison, kept out of line so that this common path stays short
    if (same_point(t.a, cp)) return cp;                       // (often: the closest point IS the first vertex)
    const double da = sqlen3(t.a - o), dc = sqlen3(cp - o);
    if (da > dc + dc * 1e-12 && dc > 1e-290) return cp;
    const int r = cmp_dist_out_of_line(o.x, o.y, o.z, t.a.x, t.a.y, t.a.z, cp.x, cp.y, cp.z);
    if (r & 4) ++nexact;
    return (r & 3) == 0 ? t.a : cp;
}

// The same for a Segment_3 primitive (AABB_segment_primitive_3 / AABB_halfedge_graph_segment_primitive; reference point = the source, AABB_segment_primitive_3.h:40-58):
// Construct_projected_point_3(Segment_3, Point_3) (function_objects.h:3266-3285), or the source when that is strictly nearer than the rounded projection.
B2_HD Vec3 closest_candidate_segment(const Vec3& a, const Vec3& b, const Vec3& o, int& nexact) {
    const Vec3 cp = closest_on_segment(a, b, o);
    if (same_point(a, cp)) return cp;
    const double da = sqlen3(a - o), dc = sqlen3(cp - o);
    if (da > dc + dc * 1e-12 && dc > 1e-290) return cp;
    const int r = cmp_dist_out_of_line(o.x, o.y, o.z, a.x, a.y, a.z, cp.x, cp.y, cp.z);
    if (r & 4) ++nexact;
    return (r & 3) == 0 ? a : cp;
}

}  // namespace b2
