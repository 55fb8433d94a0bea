// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under cgal_b200/ may include, link or call this.
//
// CPU restatement of the Epick predicates and plain-double constructions that CGAL's AABB_tree hot path
// reaches through AABB_traits_3.  Every function cites the reference file:line it
// follows (paths relative to the CGAL source root).  Constructions replicate the reference's sequence of
// IEEE double operations (compile with -ffp-contract=off); predicates return the exact sign, evaluated as
// "static filter -> exact" (the reference's middle Interval_nt stage cannot change the answer, only the
// speed, and is skipped).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <optional>
#include <utility>
#include "exact.hpp"

namespace orc {

struct P3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    bool operator==(const P3& o) const { return x == o.x && y == o.y && z == o.z; }  // Equal_3: Static_filters/Equal_3.h:44-61
    bool operator!=(const P3& o) const { return !(*this == o); }
};
typedef P3 V3;
static inline V3 vsub(const P3& a, const P3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // Vector_3(b -> a) = a - b
static inline P3 padd(const P3& a, const V3& v) { return {a.x + v.x, a.y + v.y, a.z + v.z}; }
// Construct_cross_product_vector_3: Cartesian_kernel/include/CGAL/Cartesian/function_objects.h:2263-2268
static inline V3 cross(const V3& v, const V3& w) {
    return {v.y * w.z - v.z * w.y, v.z * w.x - v.x * w.z, v.x * w.y - v.y * w.x};
}
// Compute_scalar_product_3: x*x' + y*y' + z*z', left to right
static inline double dot(const V3& v, const V3& w) { return v.x * w.x + v.y * w.y + v.z * w.z; }
// Compute_squared_distance_3(Point_3,Point_3): Kernel_23/include/CGAL/Kernel/function_objects.h:1130-1136
// -> squared length of (q - p): Cartesian_kernel/include/CGAL/Cartesian/function_objects.h:1142-1146
static inline double sqdist(const P3& p, const P3& q) {
    V3 v = vsub(q, p);
    return v.x * v.x + v.y * v.y + v.z * v.z;
}

struct Tri { P3 a, b, c; };      // Triangle_3(v0, v1, v2) in Surface_mesh/OFF order
struct Box { double lo[3], hi[3]; };  // Bbox_3

struct Stats { uint64_t orient_calls = 0, orient_exact = 0, cmpdist_calls = 0, cmpdist_exact = 0, bbox_exact = 0;
               uint64_t box_tests = 0, prim_tests = 0; };   // traversal work: child boxes tested, primitives tested (SURVEY.md §8d asks for them next to the GPU's)
extern thread_local Stats g_stats;

// ------------------------------------------------------------------------------------------------
// Orientation_3  (Filtered_kernel/include/CGAL/Filtered_kernel/internal/Static_filters/Orientation_3.h:158-255)
// exact stage = orientationC3 (Cartesian_kernel/include/CGAL/predicates/kernel_ftC3.h:134-142):
//   sign det [ q-p | r-p | s-p ]  (Kernel_23/include/CGAL/determinant.h:41-52)
// ------------------------------------------------------------------------------------------------
static inline int exact_det3_of_diffs(const Expansion& ax, const Expansion& ay, const Expansion& az,
                                      const Expansion& bx, const Expansion& by, const Expansion& bz,
                                      const Expansion& cx, const Expansion& cy, const Expansion& cz) {
    // det [a b c] with a,b,c as columns (x;y;z) == rows; cofactor expansion along c
    Expansion m01 = ax * by - ay * bx;   // xy minor
    Expansion m02 = ax * bz - az * bx;   // xz minor
    Expansion m12 = ay * bz - az * by;   // yz minor
    Expansion d = m01 * cz - m02 * cy + m12 * cx;
    return d.sign();
}

static inline int orientation_exact(const P3& p, const P3& q, const P3& r, const P3& s) {
    return exact_det3_of_diffs(ex_diff(q.x, p.x), ex_diff(q.y, p.y), ex_diff(q.z, p.z),
                               ex_diff(r.x, p.x), ex_diff(r.y, p.y), ex_diff(r.z, p.z),
                               ex_diff(s.x, p.x), ex_diff(s.y, p.y), ex_diff(s.z, p.z));
}

static inline int orientation(const P3& p, const P3& q, const P3& r, const P3& s) {
    ++g_stats.orient_calls;
    double pqx = q.x - p.x, pqy = q.y - p.y, pqz = q.z - p.z;
    double prx = r.x - p.x, pry = r.y - p.y, prz = r.z - p.z;
    double psx = s.x - p.x, psy = s.y - p.y, psz = s.z - p.z;
    double maxx = std::fabs(pqx), maxy = std::fabs(pqy), maxz = std::fabs(pqz);
    double aprx = std::fabs(prx), apsx = std::fabs(psx);
    double apry = std::fabs(pry), apsy = std::fabs(psy);
    double aprz = std::fabs(prz), apsz = std::fabs(psz);
    if (maxx < aprx) maxx = aprx;
    if (maxx < apsx) maxx = apsx;
    if (maxy < apry) maxy = apry;
    if (maxy < apsy) maxy = apsy;
    if (maxz < aprz) maxz = aprz;
    if (maxz < apsz) maxz = apsz;
    // determinant(pqx,pqy,pqz, prx,pry,prz, psx,psy,psz): determinant.h:41-52
    const double m01 = pqx * pry - prx * pqy;
    const double m02 = pqx * psy - psx * pqy;
    const double m12 = prx * psy - psx * pry;
    const double det = m01 * psz - m02 * prz + m12 * pqz;
    double eps = 5.1107127829973299e-15 * maxx * maxy * maxz;
    if (maxx > maxz) std::swap(maxx, maxz);
    if (maxy > maxz) std::swap(maxy, maxz);
    else if (maxy < maxx) std::swap(maxx, maxy);
    if (maxx < 1e-97) {
        if (maxx == 0) return 0;
    } else if (maxz < 1e102) {
        if (det > eps) return 1;
        if (det < -eps) return -1;
    }
    ++g_stats.orient_exact;
    return orientation_exact(p, q, r, s);
}

// orientation(a,b,c, a + (q-p)) with the translated point evaluated EXACTLY, as happens when
// Ray_3_Triangle_3_do_intersect.h:218-219 runs inside Filtered_predicate's exact kernel:
//   sign det [ b-a | c-a | q-p ]
static inline int orientation_dir(const P3& a, const P3& b, const P3& c, const P3& p, const P3& q) {
    ++g_stats.orient_calls;
    double bax = b.x - a.x, bay = b.y - a.y, baz = b.z - a.z;
    double cax = c.x - a.x, cay = c.y - a.y, caz = c.z - a.z;
    double dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    double maxx = std::max(std::fabs(bax), std::max(std::fabs(cax), std::fabs(dx)));
    double maxy = std::max(std::fabs(bay), std::max(std::fabs(cay), std::fabs(dy)));
    double maxz = std::max(std::fabs(baz), std::max(std::fabs(caz), std::fabs(dz)));
    const double m01 = bax * cay - cax * bay;
    const double m02 = bax * dy - dx * bay;
    const double m12 = cax * dy - dx * cay;
    const double det = m01 * dz - m02 * caz + m12 * baz;
    double eps = 5.1107127829973299e-15 * maxx * maxy * maxz;   // same error model: 9 rounded differences
    double lo = std::min(maxx, std::min(maxy, maxz)), hi = std::max(maxx, std::max(maxy, maxz));
    if (lo >= 1e-97 && hi < 1e102) {
        if (det > eps) return 1;
        if (det < -eps) return -1;
    }
    ++g_stats.orient_exact;
    return exact_det3_of_diffs(ex_diff(b.x, a.x), ex_diff(b.y, a.y), ex_diff(b.z, a.z),
                               ex_diff(c.x, a.x), ex_diff(c.y, a.y), ex_diff(c.z, a.z),
                               ex_diff(q.x, p.x), ex_diff(q.y, p.y), ex_diff(q.z, p.z));
}

// ------------------------------------------------------------------------------------------------
// orientationC2 / coplanar_orientationC3  (kernel_ftC2.h:403-408, kernel_ftC3.h:196-232).
// Coplanar_orientation_3 has no static filter in Epick (Static_filters.h:78-79,338-345): interval -> exact.
// ------------------------------------------------------------------------------------------------
static inline int orientation2(double px, double py, double qx, double qy, double rx, double ry) {
    // sign_of_determinant(qx-px, qy-py, rx-px, ry-py) = compare((qx-px)*(ry-py), (rx-px)*(qy-py))
    double a = qx - px, b = qy - py, c = rx - px, d = ry - py;
    double l = a * d, r = c * b, det = l - r;
    double err = 8.8817841970012523e-16 * (std::fabs(l) + std::fabs(r));   // 8 eps, generous
    double mag = std::max(std::max(std::fabs(a), std::fabs(b)), std::max(std::fabs(c), std::fabs(d)));
    if (mag > 1e-140 && mag < 1e140) {
        if (det > err) return 1;
        if (det < -err) return -1;
    }
    Expansion e = ex_diff(qx, px) * ex_diff(ry, py) - ex_diff(rx, px) * ex_diff(qy, py);
    return e.sign();
}
static inline int coplanar_orientation(const P3& p, const P3& q, const P3& r) {
    int oxy = orientation2(p.x, p.y, q.x, q.y, r.x, r.y);
    if (oxy != 0) return oxy;
    int oyz = orientation2(p.y, p.z, q.y, q.z, r.y, r.z);
    if (oyz != 0) return oyz;
    return orientation2(p.x, p.z, q.x, q.z, r.x, r.z);
}
static inline int coplanar_orientation(const P3& p, const P3& q, const P3& r, const P3& s) {
    int oxy = orientation2(p.x, p.y, q.x, q.y, r.x, r.y);
    if (oxy != 0) return oxy * orientation2(p.x, p.y, q.x, q.y, s.x, s.y);
    int oyz = orientation2(p.y, p.z, q.y, q.z, r.y, r.z);
    if (oyz != 0) return oyz * orientation2(p.y, p.z, q.y, q.z, s.y, s.z);
    int oxz = orientation2(p.x, p.z, q.x, q.z, r.x, r.z);
    return oxz * orientation2(p.x, p.z, q.x, q.z, s.x, s.z);
}
// collinear_are_ordered_along_lineC3: kernel_ftC3.h:279-291
static inline bool collinear_ordered(const P3& p, const P3& q, const P3& r) {
    if (p.x < q.x) return !(r.x < q.x);
    if (q.x < p.x) return !(q.x < r.x);
    if (p.y < q.y) return !(r.y < q.y);
    if (q.y < p.y) return !(q.y < r.y);
    if (p.z < q.z) return !(r.z < q.z);
    if (q.z < p.z) return !(q.z < r.z);
    return true;
}

// ------------------------------------------------------------------------------------------------
// Compare_distance_3(p,q,r) = sign(|q-p|^2 - |r-p|^2)
// (Static_filters/Compare_distance_3.h:42-134; exact: cmp_dist_to_pointC3 kernel_ftC3.h:426-432)
// ------------------------------------------------------------------------------------------------
static inline int compare_distance_exact(const P3& p, const P3& q, const P3& r) {
    Expansion acc;
    for (int i = 0; i < 3; ++i) {
        Expansion dq = ex_diff(q[i], p[i]), dr = ex_diff(r[i], p[i]);
        acc = acc + dq * dq - dr * dr;
    }
    return acc.sign();
}
static inline int compare_distance(const P3& p, const P3& q, const P3& r) {
    ++g_stats.cmpdist_calls;
    if (q == r) return 0;
    double qpx = q.x - p.x, qpy = q.y - p.y, qpz = q.z - p.z;
    double qp2 = (qpx * qpx + qpy * qpy) + qpz * qpz;
    double rpx = r.x - p.x, rpy = r.y - p.y, rpz = r.z - p.z;
    double rp2 = (rpx * rpx + rpy * rpy) + rpz * rpz;
    double diff = qp2 - rp2;
    double max1 = std::fabs(qpx);
    if (max1 < std::fabs(qpy)) max1 = std::fabs(qpy);
    if (max1 < std::fabs(qpz)) max1 = std::fabs(qpz);
    if (max1 < std::fabs(rpx)) max1 = std::fabs(rpx);
    if (max1 < std::fabs(rpy)) max1 = std::fabs(rpy);
    if (max1 < std::fabs(rpz)) max1 = std::fabs(rpz);
    if (!(max1 < 2.42701102401884262260e-147) && !(max1 > 8.37987995621411946582e+152)) {
        double eps = 3.77746921267322435884e-15 * (max1 * max1);
        if (diff > eps) return 1;
        if (diff < -eps) return -1;
    }
    ++g_stats.cmpdist_exact;
    return compare_distance_exact(p, q, r);
}

// ------------------------------------------------------------------------------------------------
// Do_intersect_3(Segment_3 | Ray_3, Bbox_3)
// Static_filters/Do_intersect_3.h:93-126 (segment: bounded_0 = bounded_1 = true), :175-208 (ray: bounded_1 = false)
// -> Intersections_3/internal/Bbox_3_Segment_3_do_intersect.h:130-448 (error model :61-128).
// Templated on the number type like the reference: NT = Filt (double + static error bound, may answer
// "uncertain") or NT = Expansion (exact).
// ------------------------------------------------------------------------------------------------
struct Unc { int v; /* 0 false, 1 true, 2 indeterminate */ };

struct GreaterFilt {      // Do_intersect_bbox_segment_aux_is_greater<double, bounded_0, true>
    bool bounded_0; double error = 0, tmax = 0, dmax = 0;
    void reg(double t, double d) {
        if (bounded_0) { if (t > tmax) tmax = t; if (d > dmax) dmax = d; }
        else { double at = std::fabs(t), ad = std::fabs(d); if (at > tmax) tmax = at; if (ad > dmax) dmax = ad; }
    }
    void bound() { error = tmax * dmax * 8.8872057372592798e-16; }
    bool bad() const { return dmax > 1e153 || tmax > 1e153 || dmax < 1e-146 || tmax < 1e-146; }
    Unc gt(double a, double b) const { double x = a - b; if (x > error) return {1}; if (x < -error) return {0}; return {2}; }
};

template <bool EXACT> struct BSNum;
template <> struct BSNum<false> {
    typedef double T;
    static T sub(double a, double b) { return a - b; }
    static T one() { return 1.0; }
    static T max0(const T& t) { return std::max(0.0, t); }
};
template <> struct BSNum<true> {
    typedef Expansion T;
    static T sub(double a, double b) { return ex_diff(a, b); }
    static T one() { return Expansion(1.0); }
    static T max0(const T& t) { return t.sign() < 0 ? Expansion() : t; }
};

template <bool EXACT>
static Unc bbox_segment_aux(const P3& p, const P3& q, const Box& bb, bool bounded_0, bool bounded_1) {
    typedef BSNum<EXACT> N; typedef typename N::T T;
    bool pin = true, qin = true;
    for (int i = 0; i < 3; ++i) {
        pin = pin && p[i] >= bb.lo[i] && p[i] <= bb.hi[i];
        qin = qin && q[i] >= bb.lo[i] && q[i] <= bb.hi[i];
    }
    if (pin || qin) return {1};
    T tmn[3], dmn[3], tmx[3], dmx[3];
    for (int i = 0; i < 3; ++i) {
        const double pi = p[i], qi = q[i], bmin = bb.lo[i], bmax = bb.hi[i];
        if (qi >= pi) {
            if (bounded_0 && pi > bmax) return {0};
            if (bounded_1 && qi < bmin) return {0};
            if (bounded_1 && bmax > qi) { tmx[i] = N::one(); dmx[i] = N::one(); }
            else { tmx[i] = N::sub(bmax, pi); dmx[i] = N::sub(qi, pi); }
            tmn[i] = N::sub(bmin, pi); dmn[i] = N::sub(qi, pi);
        } else {
            if (bounded_1 && qi > bmax) return {0};
            if (bounded_0 && pi < bmin) return {0};
            if (bounded_1 && bmin < qi) { tmx[i] = N::one(); dmx[i] = N::one(); }
            else { tmx[i] = N::sub(pi, bmin); dmx[i] = N::sub(pi, qi); }
            tmn[i] = N::sub(pi, bmax); dmn[i] = N::sub(pi, qi);
        }
        if (bounded_0) tmn[i] = N::max0(tmn[i]);
        if (pi == qi && !(bounded_0 && bounded_1)) {
            if (pi > bmax || pi < bmin) return {0};
        }
    }
    T tmin = tmn[0], dmin = dmn[0], tmax = tmx[0], dmax = dmx[0];
    GreaterFilt g; g.bounded_0 = bounded_0;
    auto gt = [&](const T& a, const T& b) -> Unc {
        if constexpr (EXACT) { return {(a - b).sign() > 0 ? 1 : 0}; }
        else { return g.gt(a, b); }
    };
    if constexpr (!EXACT) {
        g.reg(tmin, dmin); g.reg(tmn[1], dmn[1]); g.reg(tmax, dmax); g.reg(tmx[1], dmx[1]);
        g.bound();
        if (g.bad()) return {2};
    }
    const bool xflat = (p.x == q.x), yflat = (p.y == q.y), zflat = (p.z == q.z);
    if (!yflat && !xflat) {
        Unc b1 = gt(dmx[1] * tmin, dmin * tmx[1]);
        if (b1.v != 0) return {b1.v == 2 ? 2 : 0};
        Unc b2 = gt(dmax * tmn[1], dmn[1] * tmax);
        if (b2.v != 0) return {b2.v == 2 ? 2 : 0};
    }
    Unc b = {0};
    if (xflat || (!yflat && (b = gt(dmin * tmn[1], dmn[1] * tmin)).v == 1)) { tmin = tmn[1]; dmin = dmn[1]; }
    if (b.v == 2) return b;
    if (xflat || (!yflat && (b = gt(dmx[1] * tmax, dmax * tmx[1])).v == 1)) { tmax = tmx[1]; dmax = dmx[1]; }
    if (b.v == 2) return b;
    if ((!xflat || !yflat) && !zflat) {
        if constexpr (!EXACT) {
            g.reg(tmn[2], dmn[2]); g.reg(tmx[2], dmx[2]);
            g.bound();
            if (g.bad()) return {2};
        }
        Unc b1 = gt(dmx[2] * tmin, dmin * tmx[2]);
        if (b1.v != 0) return {b1.v == 2 ? 2 : 0};
        Unc b2 = gt(dmax * tmn[2], dmn[2] * tmax);
        if (b2.v != 0) return {b2.v == 2 ? 2 : 0};
    }
    return {1};
}

static inline bool do_intersect_bbox(const P3& p, const P3& q, const Box& bb, bool is_ray) {
    Unc u = bbox_segment_aux<false>(p, q, bb, true, !is_ray);
    if (u.v != 2) return u.v == 1;
    ++g_stats.bbox_exact;
    return bbox_segment_aux<true>(p, q, bb, true, !is_ray).v == 1;
}

// ------------------------------------------------------------------------------------------------
// Line_3 and Plane_3 queries.  Do_intersect_3 on (Line_3 | Plane_3, Triangle_3 | Bbox_3) has no static filter in Epick
// (Filtered_kernel/include/CGAL/Filtered_kernel/internal/Static_filters/ holds none for these pairs), so the WHOLE function body —
// including point_on(l, 1) = l.point() + l.to_vector() (Cartesian/Line_3.h) — runs inside Filtered_predicate's interval and then exact
// kernel: the answer is the exact evaluation of the decision tree on the STORED members (point + direction vector; a, b, c, d).
// The restatement evaluates every sign with expansion arithmetic and no filter.
// ------------------------------------------------------------------------------------------------
static inline Expansion ex_of(double v) { return Expansion(v); }
// sign det [ u | v | w ] of three exact vectors (columns)
static inline int exact_det3(const Expansion u[3], const Expansion v[3], const Expansion w[3]) {
    return exact_det3_of_diffs(u[0], u[1], u[2], v[0], v[1], v[2], w[0], w[1], w[2]);
}
static inline void ex_vec(const P3& to, const P3& from, Expansion out[3]) { out[0] = ex_diff(to.x, from.x); out[1] = ex_diff(to.y, from.y); out[2] = ex_diff(to.z, from.z); }
// orientation(p, p+v, r, s) = sign det [ v | r-p | s-p ]     (kernel_ftC3.h:134-142 with q - p = v exactly)
static inline int orientation_line(const P3& p, const V3& v, const P3& r, const P3& s) {
    Expansion a[3] = {ex_of(v.x), ex_of(v.y), ex_of(v.z)}, b[3], c[3];
    ex_vec(r, p, b); ex_vec(s, p, c);
    return exact_det3(a, b, c);
}
// orientation(a, b, c, p + v) = sign det [ b-a | c-a | (p-a)+v ]
static inline int orientation_translated(const P3& a, const P3& b, const P3& c, const P3& p, const V3& v) {
    Expansion u[3], w[3], x[3];
    ex_vec(b, a, u); ex_vec(c, a, w); ex_vec(p, a, x);
    x[0] = x[0] + ex_of(v.x); x[1] = x[1] + ex_of(v.y); x[2] = x[2] + ex_of(v.z);
    return exact_det3(u, w, x);
}
// coplanar_orientation(p, p+v, r): coplanar_orientationC3 (kernel_ftC3.h:196-211) with orientationC2(p, q, r) = sign(vx (ry-py) - vy (rx-px))
static inline int coplanar_orientation_line(const P3& p, const V3& v, const P3& r) {
    auto o2 = [](double va, double vb, double pa, double pb, double ra, double rb) { return (ex_of(va) * ex_diff(rb, pb) - ex_of(vb) * ex_diff(ra, pa)).sign(); };
    int oxy = o2(v.x, v.y, p.x, p.y, r.x, r.y);
    if (oxy != 0) return oxy;
    int oyz = o2(v.y, v.z, p.y, p.z, r.y, r.z);
    if (oyz != 0) return oyz;
    return o2(v.x, v.z, p.x, p.z, r.x, r.z);
}
// Do_intersect_3(Triangle_3, Line_3): Intersections_3/include/CGAL/Intersections_3/internal/Line_3_Triangle_3_do_intersect.h:27-79
static inline bool do_intersect_line_tri(const Tri& t, const P3& p, const V3& v) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    if (orientation(a, b, c, p) != 0 || orientation_translated(a, b, c, p, v) != 0) {
        const int pqab = orientation_line(p, v, a, b), pqbc = orientation_line(p, v, b, c);
        switch (pqab) {
            case 1: return pqbc != -1 && orientation_line(p, v, c, a) != -1;
            case -1: return pqbc != 1 && orientation_line(p, v, c, a) != 1;
            default:
                if (pqbc == 1) return orientation_line(p, v, c, a) != -1;
                if (pqbc == -1) return orientation_line(p, v, c, a) != 1;
                return true;
        }
    }
    const int pqa = coplanar_orientation_line(p, v, a);
    return coplanar_orientation_line(p, v, b) != pqa || coplanar_orientation_line(p, v, c) != pqa;
}
// Do_intersect_3(Line_3, Bbox_3): Bbox_3_Line_3_do_intersect.h:28-131, every comparison exact
static inline bool do_intersect_line_bbox(const P3& p, const V3& v, const Box& bb) {
    const double pc[3] = {p.x, p.y, p.z}, vc[3] = {v.x, v.y, v.z};
    if (pc[0] >= bb.lo[0] && pc[0] <= bb.hi[0] && pc[1] >= bb.lo[1] && pc[1] <= bb.hi[1] && pc[2] >= bb.lo[2] && pc[2] <= bb.hi[2]) return true;
    auto axis = [&](int k, Expansion& tmin, Expansion& tmax, Expansion& d) {
        if (vc[k] < 0) { tmin = ex_diff(pc[k], bb.hi[k]); tmax = ex_diff(pc[k], bb.lo[k]); d = ex_of(-vc[k]); }
        else { tmin = ex_diff(bb.lo[k], pc[k]); tmax = ex_diff(bb.hi[k], pc[k]); d = ex_of(vc[k]); }
    };
    auto less = [](const Expansion& x, const Expansion& y) { return (x - y).sign() < 0; };
    Expansion dmin, tmin, tmax, d_, tmin_, tmax_;
    axis(0, tmin, tmax, dmin);
    if (dmin.sign() == 0 && (tmin.sign() > 0 || tmax.sign() < 0)) return false;
    Expansion dmax = dmin;
    axis(1, tmin_, tmax_, d_);
    if (d_.sign() == 0) {
        if (tmin_.sign() > 0 || tmax_.sign() < 0) return false;
    } else {
        if (less(dmin * tmax_, d_ * tmin) || less(d_ * tmax, dmax * tmin_)) return false;
    }
    if (less(d_ * tmin, dmin * tmin_)) { tmin = tmin_; dmin = d_; }
    if (less(dmax * tmax_, d_ * tmax)) { tmax = tmax_; dmax = d_; }
    axis(2, tmin_, tmax_, d_);
    return !less(dmin * tmax_, d_ * tmin) && !less(d_ * tmax, dmax * tmin_);
}
// Oriented_side_3(Plane_3, Point_3) = side_of_oriented_planeC3 (kernel_ftC3.h:402-409): sign(a px + b py + c pz + d)
static inline int plane_oriented_side(const double h[4], const P3& p) {
    return (ex_of(h[0]) * ex_of(p.x) + ex_of(h[1]) * ex_of(p.y) + ex_of(h[2]) * ex_of(p.z) + ex_of(h[3])).sign();
}
// Do_intersect_3(Triangle_3, Plane_3): Plane_3_Triangle_3_do_intersect.h:27-51
static inline bool do_intersect_plane_tri(const Tri& t, const double h[4]) {
    const int s0 = plane_oriented_side(h, t.a);
    if (s0 == 0) return true;
    return plane_oriented_side(h, t.b) != s0 || plane_oriented_side(h, t.c) != s0;
}
// Do_intersect_3(Plane_3, Bbox_3): Iso_cuboid_3_Plane_3_do_intersect.h:28-125 (get_min_max by the signs of the normal, certain in the
// exact kernel; the eight-corner fallback of the interval stage decides the same exact fact)
static inline bool do_intersect_plane_bbox(const double h[4], const Box& bb) {
    P3 pmin, pmax;
    pmin.x = h[0] > 0 ? bb.lo[0] : bb.hi[0]; pmax.x = h[0] > 0 ? bb.hi[0] : bb.lo[0];
    pmin.y = h[1] > 0 ? bb.lo[1] : bb.hi[1]; pmax.y = h[1] > 0 ? bb.hi[1] : bb.lo[1];
    pmin.z = h[2] > 0 ? bb.lo[2] : bb.hi[2]; pmax.z = h[2] > 0 ? bb.hi[2] : bb.lo[2];
    return !(plane_oriented_side(h, pmax) < 0 || plane_oriented_side(h, pmin) > 0);
}

// ------------------------------------------------------------------------------------------------
// Do_intersect_3(Segment_3, Triangle_3): Intersections_3/internal/Segment_3_Triangle_3_do_intersect.h:24-152,177-262
// ------------------------------------------------------------------------------------------------
static inline bool seg_tri_coplanar(const P3& A, const P3& B, const P3& C, const P3& p, const P3& q) {
    const P3* a = &A; const P3* b = &B; const P3* c = &C;
    if (coplanar_orientation(A, B, C) != 1) { b = &C; c = &B; }
    const int pqa = coplanar_orientation(p, q, *a), pqb = coplanar_orientation(p, q, *b), pqc = coplanar_orientation(p, q, *c);
    auto co = [](const P3& x, const P3& y, const P3& z) { return coplanar_orientation(x, y, z); };
    switch (pqa) {
    case 1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return false;
            return co(*b, *c, q) != -1 && co(*c, *a, p) != -1;
        case -1:
        case 0:
            if (pqc == 1) return co(*a, *b, q) != -1 && co(*b, *c, p) != -1;
            return co(*a, *b, q) != -1 && co(*c, *a, p) != -1;
        }
        return false;
    case -1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return co(*a, *b, p) != -1 && co(*c, *a, q) != -1;
            return co(*a, *b, p) != -1 && co(*b, *c, q) != -1;
        case -1:
            if (pqc == -1) return false;
            return co(*b, *c, p) != -1 && co(*c, *a, q) != -1;
        case 0:
            if (pqc == -1) return co(*a, *b, p) != -1 && co(*b, *c, q) != -1;
            return co(*a, *b, p) != -1 && co(*c, *a, q) != -1;
        }
        return false;
    default:  // COLLINEAR
        switch (pqb) {
        case 1:
            if (pqc == 1) return co(*a, *b, p) != -1 && co(*c, *a, q) != -1;
            return co(*a, *b, p) != -1 && co(*b, *c, q) != -1;
        case -1:
            if (pqc == -1) return co(*a, *b, q) != -1 && co(*c, *a, p) != -1;
            return co(*a, *b, q) != -1 && co(*b, *c, p) != -1;
        default:
            if (pqc == 1) return co(*b, *c, p) != -1 && co(*c, *a, q) != -1;
            return co(*b, *c, q) != -1 && co(*c, *a, p) != -1;
        }
    }
}

static inline bool do_intersect_seg_tri(const Tri& t, const P3& p, const P3& q) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    const int abcp = orientation(a, b, c, p), abcq = orientation(a, b, c, q);   // the dual 5-point call, :193
    switch (abcp) {
    case 1:
        if (abcq == 1) return false;
        return orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1;
    case -1:
        if (abcq == -1) return false;
        return orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1;
    default:
        switch (abcq) {
        case 1: return orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1;
        case -1: return orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1;
        default: return seg_tri_coplanar(a, b, c, p, q);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Do_intersect_3(Triangle_3, Triangle_3): Intersections_3/internal/Triangle_3_Triangle_3_do_intersect.h (Guigue & Devillers):
// _intersection_test_vertex :24-67, _intersection_test_edge :69-98, do_intersect_coplanar :100-152, do_intersect :154-450.
// ------------------------------------------------------------------------------------------------
static inline bool tt_test_vertex(const P3* p, const P3* q, const P3* r, const P3* a, const P3* b, const P3* c) {   // "vertex p sees vertex c"
    auto co = [](const P3& x, const P3& y, const P3& z) { return coplanar_orientation(x, y, z); };
    if (co(*c, *a, *q) != -1) {
        if (co(*c, *b, *q) != 1) {
            if (co(*p, *a, *q) == 1) return co(*p, *b, *q) != 1;
            return co(*p, *a, *r) != -1 && co(*q, *r, *a) != -1;
        }
        if (co(*p, *b, *q) != 1) return co(*c, *b, *r) != 1 && co(*q, *r, *b) != -1;
        return false;
    }
    if (co(*c, *a, *r) != -1) {                                                  // qr straddles (ac)
        if (co(*q, *r, *c) != -1) return co(*p, *a, *r) != -1;
        return co(*q, *r, *b) != -1 && co(*c, *r, *b) != -1;
    }
    return false;                                                                 // ca separates
}
static inline bool tt_test_edge(const P3* p, const P3* q, const P3* r, const P3* a, const P3* /*b*/, const P3* c) {   // "vertex p sees edge ca"
    auto co = [](const P3& x, const P3& y, const P3& z) { return coplanar_orientation(x, y, z); };
    if (co(*c, *a, *q) != -1) {                                                  // pq straddles (ac)
        if (co(*p, *a, *q) != -1) return co(*p, *q, *c) != -1;
        return co(*q, *r, *a) != -1 && co(*r, *p, *a) != -1;
    }
    if (co(*c, *a, *r) != -1) return co(*p, *a, *r) != -1 && (co(*p, *r, *c) != -1 || co(*q, *r, *c) != -1);
    return false;                                                                 // ac separates
}
static inline bool tri_tri_coplanar(const Tri& t1, const Tri& t2) {
    auto co = [](const P3& x, const P3& y, const P3& z) { return coplanar_orientation(x, y, z); };
    const P3 *p = &t1.a, *q = &t1.b, *r = &t1.c, *a = &t2.a, *b = &t2.b, *c = &t2.c;
    if (co(t1.a, t1.b, t1.c) == -1) { q = &t1.c; r = &t1.b; }                     // both counterclockwise (:120-130)
    if (co(t2.a, t2.b, t2.c) == -1) { b = &t2.c; c = &t2.b; }
    if (co(*a, *b, *p) != -1) {                                                   // p among the lines supporting abc's edges (:135-151)
        if (co(*b, *c, *p) != -1) {
            if (co(*c, *a, *p) != -1) return true;                                // p inside abc
            return tt_test_edge(p, q, r, a, b, c);                                // p sees ac
        }
        if (co(*c, *a, *p) != -1) return tt_test_edge(p, q, r, c, a, b);          // p sees bc
        return tt_test_vertex(p, q, r, a, b, c);                                  // p sees c
    }
    if (co(*b, *c, *p) != -1) {
        if (co(*c, *a, *p) != -1) return tt_test_edge(p, q, r, b, c, a);          // p sees ab
        return tt_test_vertex(p, q, r, b, c, a);                                  // p sees a
    }
    return tt_test_vertex(p, q, r, c, a, b);                                      // p sees b
}
// the case analysis of :180-302 (and again :316-440 for the second triangle): which two edges of (p, q, r) cross the other triangle's plane, given the
// orientations dp, dq, dr of its vertices; returns 0 = all strictly on one side, 2 = all coplanar, 1 = edges (s_min, t_min), (s_max, t_max) set
static inline int tt_pick_edges(int dp, int dq, int dr, const P3* p, const P3* q, const P3* r, const P3*& s_min, const P3*& t_min, const P3*& s_max, const P3*& t_max) {
    auto set = [&](const P3* a, const P3* b, const P3* c, const P3* d) { s_min = a; t_min = b; s_max = c; t_max = d; return 1; };
    switch (dp) {
    case 1:
        if (dq == 1) { if (dr == 1) return 0; return set(q, r, r, p); }           // r isolated on the negative side
        if (dr == 1) return set(p, q, q, r);                                      // q isolated on the negative side
        return set(p, q, r, p);                                                   // p isolated on the positive side
    case -1:
        if (dq == -1) { if (dr == -1) return 0; return set(r, p, q, r); }         // r isolated on the positive side
        if (dr == -1) return set(q, r, p, q);                                     // q isolated on the positive side
        return set(r, p, p, q);                                                   // p isolated on the negative side
    default:
        switch (dq) {
        case 1: return dr == 1 ? set(r, p, p, q) : set(q, r, p, q);
        case -1: return dr == -1 ? set(p, q, r, p) : set(p, q, q, r);
        default:
            if (dr == 1) return set(r, p, q, r);
            if (dr == -1) return set(q, r, r, p);
            return 2;
        }
    }
}
// t1 = query, t2 = primitive (AABB_traits_3.h:673-677 calls do_intersect_3_object()(q, datum))
static inline bool do_intersect_tri_tri(const Tri& t1, const Tri& t2) {
    const P3 &p = t1.a, &q = t1.b, &r = t1.c, &a = t2.a, &b = t2.b, &c = t2.c;
    const P3 *s_min1, *t_min1, *s_max1, *t_max1, *s_min2, *t_min2, *s_max2, *t_max2;
    const int dp = orientation(a, b, c, p), dq = orientation(a, b, c, q), dr = orientation(a, b, c, r);      // :176-178
    switch (tt_pick_edges(dp, dq, dr, &p, &q, &r, s_min1, t_min1, s_max1, t_max1)) {
    case 0: return false;
    case 2: return tri_tri_coplanar(t1, t2);
    default: break;
    }
    const int da = orientation(p, q, r, a), db = orientation(p, q, r, b), dc = orientation(p, q, r, c);      // :312-314
    switch (tt_pick_edges(da, db, dc, &a, &b, &c, s_min2, t_min2, s_max2, t_max2)) {
    case 0: return false;
    case 2: return tri_tri_coplanar(t1, t2);                                      // ("supposed unreachable": non-flat triangles, :424-428)
    default: break;
    }
    return orientation(*s_min1, *t_min1, *s_min2, *t_min2) != 1 && orientation(*s_max1, *t_max1, *t_max2, *s_max2) != 1;   // :446-447
}

// ------------------------------------------------------------------------------------------------
// Do_intersect_3(Ray_3, Triangle_3): Intersections_3/internal/Ray_3_Triangle_3_do_intersect.h:75-191,193-317
// (no static-filter overload in Epick; the whole functor runs in the exact kernel when intervals fail,
//  so ray_direction is the exact sign of det[b-a, c-a, q-p]).
// ------------------------------------------------------------------------------------------------
static inline bool ray_tri_coplanar(const P3& A, const P3& B, const P3& C, const P3& p, const P3& q) {
    const P3* a = &A; const P3* b = &B; const P3* c = &C;
    if (coplanar_orientation(A, B, C) != 1) { b = &C; c = &B; }
    const int pqa = coplanar_orientation(p, q, *a), pqb = coplanar_orientation(p, q, *b), pqc = coplanar_orientation(p, q, *c);
    auto co = [](const P3& x, const P3& y, const P3& z) { return coplanar_orientation(x, y, z); };
    switch (pqa) {
    case 1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return false;
            return co(*a, *c, p) != 1;
        default:  // NEGATIVE, COLLINEAR
            if (pqc == 1) return co(*c, *b, p) != 1;
            return co(*a, *c, p) != 1;
        }
    case -1:
        switch (pqb) {
        case 1: return co(*b, *a, p) != 1;
        case -1:
            if (pqc == -1) return false;
            return co(*c, *b, p) != 1;
        default: return co(*b, *a, p) != 1;
        }
    default:
        switch (pqb) {
        case 1: return co(*b, *a, p) != 1;
        case -1:
            if (pqc == -1) return co(*a, *c, p) != 1;
            return co(*c, *b, p) != 1;
        default:
            if (pqc == 1) return co(*c, *b, p) != 1;
            return co(*a, *c, p) != 1;
        }
    }
}

static inline bool do_intersect_ray_tri(const Tri& t, const P3& p, const P3& q) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    const int dir = orientation_dir(a, b, c, p, q);
    if (dir == 0) {
        if (orientation(a, b, c, p) == 0) return ray_tri_coplanar(a, b, c, p, q);
        return false;
    }
    const int abcp = orientation(a, b, c, p);
    switch (abcp) {
    case 1:
        if (dir == 1) return false;
        return orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1;
    case -1:
        if (dir == -1) return false;
        return orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1;
    default:
        if (dir == 1) return orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1;
        return orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1;
    }
}

// The same with r3t3_do_intersect_endpoint_position_visitor (Ray_3_Triangle_3_do_intersect.h:23-62,193-317): -1 = no intersection, else the
// R3T3_intersection::type: 0 CROSS_FACET, 1 CROSS_SEGMENT, 2 CROSS_VERTEX (visitor.update counts the vanishing edge orientations), 3 COPLANAR_RAY,
// 4 ENDPOINT_IN_TRIANGLE.  Used by Side_of_triangle_mesh's generic-ray traversal traits.
static inline int classify_ray_tri(const Tri& t, const P3& p, const P3& q) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    const int dir = orientation_dir(a, b, c, p, q);
    if (dir == 0) {
        if (orientation(a, b, c, p) == 0) return ray_tri_coplanar(a, b, c, p, q) ? 3 : -1;      // visitor.ray_coplanar()
        return -1;
    }
    const int abcp = orientation(a, b, c, p);
    auto straddle = [&](const P3& x, const P3& y) -> int {
        int zeros = 0, o;
        o = orientation(x, y, a, b); if (o == 1) return -1; zeros += o == 0;
        o = orientation(x, y, b, c); if (o == 1) return -1; zeros += o == 0;
        o = orientation(x, y, c, a); if (o == 1) return -1; zeros += o == 0;
        return zeros;
    };
    switch (abcp) {
    case 1: return dir == 1 ? -1 : straddle(p, q);
    case -1: return dir == -1 ? -1 : straddle(q, p);
    default: {                                                                                     // visitor.end_point_in_triangle()
        const bool hit = dir == 1 ? (orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1)
                                  : (orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1);
        return hit ? 4 : -1;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// Intersect_3(Ray_3, Triangle_3): Intersections_3/internal/Ray_3_Triangle_3_intersection.h:28-539
// Result: kind 0 = none, 1 = Point_3 (pt[0]), 2 = Segment_3 (pt[0], pt[1]).
// ------------------------------------------------------------------------------------------------
struct RayHit { int kind = 0; P3 pt[2]; };

// plane_from_pointsC3: Cartesian_kernel/include/CGAL/constructions/kernel_ftC3.h:225-241
struct Plane { double a, b, c, d; };
static inline Plane plane_from_points(const P3& p, const P3& q, const P3& r) {
    double rpx = p.x - r.x, rpy = p.y - r.y, rpz = p.z - r.z;
    double rqx = q.x - r.x, rqy = q.y - r.y, rqz = q.z - r.z;
    Plane h;
    h.a = rpy * rqz - rqy * rpz;
    h.b = rpz * rqx - rqz * rpx;
    h.c = rpx * rqy - rqx * rpy;
    h.d = -h.a * r.x - h.b * r.y - h.c * r.z;
    return h;
}

// t3r3_intersection_aux :398-415 -> Line_3 x Plane_3 (Line_3_Plane_3_intersection.h:52-85); Point_3(hx,hy,hz,hw)
// divides iff hw != 1 (Cartesian/Vector_3.h:73-75)
static inline bool line_plane_point(const Tri& t, const P3& s, const P3& q, P3& out) {
    const Plane h = plane_from_points(t.a, t.b, t.c);
    const V3 d = vsub(q, s);    // Construct_line_3(Ray_3) = (source, second_point - source): Cartesian/function_objects.h:2706-2707
    const double num = h.a * s.x + h.b * s.y + h.c * s.z + h.d;
    const double den = h.a * d.x + h.b * d.y + h.c * d.z;
    if (den == 0) return false;
    const double hx = den * s.x - num * d.x, hy = den * s.y - num * d.y, hz = den * s.z - num * d.z;
    if (den != 1.0) out = {hx / den, hy / den, hz / den};
    else out = {hx, hy, hz};
    return true;
}

// :28-61  intersection of line (p,v) with line (a,b), all coplanar
static inline P3 coplanar_line_line(const P3& p, const V3& v, const P3& a, const P3& b) {
    const V3 ab = vsub(b, a), pa = vsub(a, p);
    const V3 pa_ab = cross(pa, ab), v_ab = cross(v, ab);
    const double t = dot(pa_ab, v_ab) / dot(v_ab, v_ab);   // sq_length = x*x + y*y + z*z
    return {p.x + v.x * t, p.y + v.y * t, p.z + v.z * t};   // trans(p, sc(v,t)); Construct_scaled_vector_3 = (c*x ...) commutative
}

// :63-146
static inline RayHit coplanar_clip(const P3& a, const P3& b, const P3& c, const P3& p, const P3& q, bool negative_side) {
    RayHit r;
    const int cap = negative_side ? coplanar_orientation(c, a, p) : coplanar_orientation(b, c, p);
    if (cap == -1) return r;
    if (cap == 0) { r.kind = 1; r.pt[0] = p; return r; }
    const V3 v = vsub(q, p);
    P3 p_end = p, q_end;
    if (negative_side) {
        if (coplanar_orientation(b, c, p) == -1) p_end = coplanar_line_line(p, v, b, c);
        q_end = coplanar_line_line(p, v, c, a);
    } else {
        if (coplanar_orientation(c, a, p) == -1) p_end = coplanar_line_line(p, v, c, a);
        q_end = coplanar_line_line(p, v, b, c);
    }
    r.kind = 2; r.pt[0] = p_end; r.pt[1] = q_end;
    return r;
}

// :147-396
static inline RayHit ray_tri_intersection_coplanar(const Tri& t, const P3& p, const P3& q) {
    const P3& a = t.a;
    const bool ccw = coplanar_orientation(t.a, t.b, t.c) == 1;
    const P3& b = ccw ? t.b : t.c;
    const P3& c = ccw ? t.c : t.b;
    const int pqa = coplanar_orientation(p, q, a), pqb = coplanar_orientation(p, q, b), pqc = coplanar_orientation(p, q, c);
    auto none = []() { return RayHit(); };
    auto point = [](const P3& x) { RayHit r; r.kind = 1; r.pt[0] = x; return r; };
    auto seg = [](const P3& x, const P3& y) { RayHit r; r.kind = 2; r.pt[0] = x; r.pt[1] = y; return r; };
    auto vertex = [&](const P3& x) { return (collinear_ordered(p, x, q) || collinear_ordered(p, q, x)) ? point(x) : none(); };
    switch (pqa) {
    case 1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return none();
            if (pqc == -1) return coplanar_clip(a, b, c, p, q, true);
            return vertex(c);
        case -1:
            if (pqc == 1) return coplanar_clip(c, a, b, p, q, true);
            return coplanar_clip(b, c, a, p, q, false);
        default:
            if (pqc == 1) return vertex(b);
            if (pqc == -1) return coplanar_clip(b, c, a, p, q, false);
            return collinear_ordered(p, b, c) ? seg(b, c) : seg(p, c);
        }
    case -1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return coplanar_clip(b, c, a, p, q, true);
            return coplanar_clip(c, a, b, p, q, false);
        case -1:
            if (pqc == 1) return coplanar_clip(a, b, c, p, q, false);
            if (pqc == -1) return none();
            return vertex(c);
        default:
            if (pqc == 1) return coplanar_clip(b, c, a, p, q, true);
            if (pqc == -1) return vertex(b);
            return collinear_ordered(p, c, b) ? seg(c, b) : seg(p, b);
        }
    default:
        switch (pqb) {
        case 1:
            if (pqc == 1) return vertex(a);
            if (pqc == -1) return coplanar_clip(c, a, b, p, q, false);
            return collinear_ordered(p, c, a) ? seg(c, a) : seg(p, a);
        case -1:
            if (pqc == 1) return coplanar_clip(c, a, b, p, q, true);
            if (pqc == -1) return vertex(a);
            return collinear_ordered(p, a, c) ? seg(a, c) : seg(p, c);
        default:
            if (pqc == 1) return collinear_ordered(p, a, b) ? seg(a, b) : seg(p, b);
            if (pqc == -1) return collinear_ordered(p, b, a) ? seg(b, a) : seg(p, a);
            return none();   // flat triangle: reference calls CGAL_error()
        }
    }
}

// :417-539
static inline RayHit ray_tri_intersection(const Tri& t, const P3& p, const P3& q) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    const P3 d = padd(a, vsub(q, p));      // a ROUNDED point: Ray_3(a, vector).second_point() = a + v  (Cartesian/Ray_3.h:48-49)
    const int dir = orientation(a, b, c, d);
    const int abcp = orientation(a, b, c, p);
    RayHit r;
    bool through;
    switch (abcp) {
    case 1:
        if (dir != -1) return r;
        through = orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1;
        break;
    case -1:
        if (dir != 1) return r;
        through = orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1;
        break;
    default:
        if (dir == 1) through = orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1;
        else if (dir == -1) through = orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1;
        else return ray_tri_intersection_coplanar(t, p, q);
    }
    if (!through) return r;
    P3 x;
    if (line_plane_point(t, p, q, x)) { r.kind = 1; r.pt[0] = x; }
    return r;
}

// ------------------------------------------------------------------------------------------------
// Intersect_3(Segment_3, Triangle_3): Intersections_3/internal/Segment_3_Triangle_3_intersection.h:28-569.  Same result record as the ray's
// (kind 0 none, 1 Point_3, 2 Segment_3).  Predicates are the filtered (exact) ones, constructions plain doubles in the reference's order.
// ------------------------------------------------------------------------------------------------
// :28-61  intersection point of the coplanar segments [p,q] and [a,b]
static inline P3 t3s3_point(const P3& p, const P3& q, const P3& a, const P3& b) { return coplanar_line_line(p, vsub(q, p), a, b); }
// :63-133  clip [p,q] into the triangle; c is isolated on the negative side of (p,q)
static inline RayHit t3s3_clip(const P3& a, const P3& b, const P3& c, const P3& p, const P3& q, bool negative_side) {
    RayHit r;
    const int bcq = coplanar_orientation(b, c, q), cap = coplanar_orientation(c, a, p);
    if (bcq == -1 || cap == -1) return r;
    if (bcq == 0) { r.kind = 1; r.pt[0] = q; return r; }
    if (cap == 0) { r.kind = 1; r.pt[0] = p; return r; }
    P3 pe = p, qe = q;
    if (coplanar_orientation(b, c, p) == -1) pe = t3s3_point(p, q, b, c);
    if (coplanar_orientation(c, a, q) == -1) qe = t3s3_point(p, q, c, a);
    r.kind = 2;
    r.pt[0] = negative_side ? pe : qe; r.pt[1] = negative_side ? qe : pe;
    return r;
}
// :135-186  [a,b] and [p,q] collinear with the same direction
static inline RayHit t3s3_collinear(const P3& a, const P3& b, const P3& p, const P3& q) {
    RayHit r;
    auto point = [&](const P3& x) { r.kind = 1; r.pt[0] = x; return r; };
    auto segment = [&](const P3& x, const P3& y) { r.kind = 2; r.pt[0] = x; r.pt[1] = y; return r; };
    if (collinear_ordered(p, a, b)) {
        if (collinear_ordered(a, b, q)) return segment(a, b);
        if (collinear_ordered(q, a, b)) return a == q ? point(a) : r;
        return segment(a, q);
    }
    if (collinear_ordered(p, b, q)) return p == b ? point(p) : segment(p, b);
    if (collinear_ordered(a, b, p)) return p == b ? point(p) : r;
    return segment(p, q);
}
// :188-398
static inline RayHit seg_tri_intersection_coplanar(const Tri& t, const P3& p, const P3& q) {
    const P3* v[3] = {&t.a, &t.b, &t.c};
    if (coplanar_orientation(t.a, t.b, t.c) != 1) std::swap(v[1], v[2]);
    const P3 &a = *v[0], &b = *v[1], &c = *v[2];
    const int pqa = coplanar_orientation(p, q, a), pqb = coplanar_orientation(p, q, b), pqc = coplanar_orientation(p, q, c);
    RayHit none;
    auto vertex_if_inside = [&](const P3& x) { RayHit r; if (collinear_ordered(p, x, q)) { r.kind = 1; r.pt[0] = x; } return r; };
    switch (pqa) {
    case 1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return none;
            if (pqc == -1) return t3s3_clip(a, b, c, p, q, true);
            return vertex_if_inside(c);
        case -1:
            if (pqc == 1) return t3s3_clip(c, a, b, p, q, true);
            return t3s3_clip(b, c, a, q, p, false);
        default:
            if (pqc == 1) return vertex_if_inside(b);
            if (pqc == -1) return t3s3_clip(b, c, a, q, p, false);
            return t3s3_collinear(b, c, p, q);
        }
    case -1:
        switch (pqb) {
        case 1:
            if (pqc == 1) return t3s3_clip(b, c, a, p, q, true);
            return t3s3_clip(c, a, b, q, p, false);
        case -1:
            if (pqc == 1) return t3s3_clip(a, b, c, q, p, false);
            if (pqc == -1) return none;
            return vertex_if_inside(c);
        default:
            if (pqc == 1) return t3s3_clip(b, c, a, p, q, true);
            if (pqc == -1) return vertex_if_inside(b);
            return t3s3_collinear(c, b, p, q);
        }
    default:
        switch (pqb) {
        case 1:
            if (pqc == 1) return vertex_if_inside(a);
            if (pqc == -1) return t3s3_clip(c, a, b, q, p, false);
            return t3s3_collinear(c, a, p, q);
        case -1:
            if (pqc == 1) return t3s3_clip(c, a, b, p, q, true);
            if (pqc == -1) return vertex_if_inside(a);
            return t3s3_collinear(a, c, p, q);
        default:
            if (pqc == 1) return t3s3_collinear(a, b, p, q);
            if (pqc == -1) return t3s3_collinear(b, a, p, q);
            return none;                                                   // (degenerate triangle: CGAL_error in the reference)
        }
    }
}
// :400-548
static inline RayHit seg_tri_intersection(const Tri& t, const P3& p, const P3& q) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    const int abcp = orientation(a, b, c, p), abcq = orientation(a, b, c, q);
    RayHit r;
    auto from_p = [&]() { return orientation(p, q, a, b) != 1 && orientation(p, q, b, c) != 1 && orientation(p, q, c, a) != 1; };
    auto from_q = [&]() { return orientation(q, p, a, b) != 1 && orientation(q, p, b, c) != 1 && orientation(q, p, c, a) != 1; };
    auto crossing = [&]() { P3 x; if (line_plane_point(t, p, q, x)) { r.kind = 1; r.pt[0] = x; } return r; };
    auto point = [&](const P3& x) { r.kind = 1; r.pt[0] = x; return r; };
    switch (abcp) {
    case 1:
        if (abcq == 1) return r;
        if (abcq == -1) return from_p() ? crossing() : r;
        return from_p() ? point(q) : r;
    case -1:
        if (abcq == 1) return from_q() ? crossing() : r;
        if (abcq == -1) return r;
        return from_q() ? point(q) : r;
    default:
        if (abcq == 1) return from_q() ? point(p) : r;
        if (abcq == -1) return from_p() ? point(p) : r;
        return seg_tri_intersection_coplanar(t, p, q);
    }
}

// ------------------------------------------------------------------------------------------------
// Intersect_3(Line_3, Triangle_3): Intersections_3/internal/Line_3_Triangle_3_intersection.h:27-419, and Intersect_3(Plane_3, Triangle_3):
// Plane_3_Triangle_3_intersection.h:27-146.  These are CONSTRUCTIONS: unlike Do_intersect_3 they run in plain double arithmetic with filtered
// predicates on the points they form, so point_on(l, 1) is the ROUNDED point p + v here (Cartesian/Line_3.h).
// ------------------------------------------------------------------------------------------------
// Intersect_3(Plane_3, Line_3(point s, direction d)): Line_3_Plane_3_intersection.h:52-85 (den == 0: no Point_3)
static inline bool plane_line_point(const double h[4], const P3& s, const V3& d, P3& out) {
    const double num = h[0] * s.x + h[1] * s.y + h[2] * s.z + h[3];
    const double den = h[0] * d.x + h[1] * d.y + h[2] * d.z;
    if (den == 0) return false;
    const double hx = den * s.x - num * d.x, hy = den * s.y - num * d.y, hz = den * s.z - num * d.z;
    if (den != 1.0) out = {hx / den, hy / den, hz / den};
    else out = {hx, hy, hz};
    return true;
}
// Line_3_Triangle_3_intersection.h:66-88  the segment cut out of the line by the edges (b,c) and (c,a)
static inline RayHit t3l3_clip(const P3& a, const P3& b, const P3& c, const P3& p, const V3& v, bool negative_side) {
    const P3 l_bc = coplanar_line_line(p, v, b, c), l_ca = coplanar_line_line(p, v, c, a);
    RayHit r; r.kind = 2;
    r.pt[0] = negative_side ? l_bc : l_ca; r.pt[1] = negative_side ? l_ca : l_bc;
    return r;
}
// :90-277
static inline RayHit line_tri_intersection_coplanar(const Tri& t, const P3& p, const P3& q, const V3& v) {
    const P3* w[3] = {&t.a, &t.b, &t.c};
    if (coplanar_orientation(t.a, t.b, t.c) != 1) std::swap(w[1], w[2]);
    const P3 &a = *w[0], &b = *w[1], &c = *w[2];
    const int pqa = coplanar_orientation(p, q, a), pqb = coplanar_orientation(p, q, b), pqc = coplanar_orientation(p, q, c);
    RayHit none;
    auto point = [](const P3& x) { RayHit r; r.kind = 1; r.pt[0] = x; return r; };
    auto segment = [](const P3& x, const P3& y) { RayHit r; r.kind = 2; r.pt[0] = x; r.pt[1] = y; return r; };
    switch (pqa) {
    case 1:
        switch (pqb) {
        case 1: return pqc == 1 ? none : (pqc == -1 ? t3l3_clip(a, b, c, p, v, true) : point(c));
        case -1: return pqc == 1 ? t3l3_clip(c, a, b, p, v, true) : t3l3_clip(b, c, a, p, v, false);
        default: return pqc == 1 ? point(b) : (pqc == -1 ? t3l3_clip(b, c, a, p, v, false) : segment(b, c));
        }
    case -1:
        switch (pqb) {
        case 1: return pqc == 1 ? t3l3_clip(b, c, a, p, v, true) : t3l3_clip(c, a, b, p, v, false);
        case -1: return pqc == 1 ? t3l3_clip(a, b, c, p, v, false) : (pqc == -1 ? none : point(c));
        default: return pqc == 1 ? t3l3_clip(b, c, a, p, v, true) : (pqc == -1 ? point(b) : segment(c, b));
        }
    default:
        switch (pqb) {
        case 1: return pqc == 1 ? point(a) : (pqc == -1 ? t3l3_clip(c, a, b, p, v, false) : segment(c, a));
        case -1: return pqc == 1 ? t3l3_clip(c, a, b, p, v, true) : (pqc == -1 ? point(a) : segment(a, c));
        default: return pqc == 1 ? segment(a, b) : (pqc == -1 ? segment(b, a) : none);
        }
    }
}
// :279-400  the line is (point p, direction vector v) as stored
static inline RayHit line_tri_intersection(const Tri& t, const P3& p, const V3& v) {
    const P3 &a = t.a, &b = t.b, &c = t.c;
    const P3 q = padd(p, v);                                   // point_on(l, 1): rounded
    RayHit none;
    auto crossing = [&]() {
        const Plane h = plane_from_points(a, b, c);
        const double hc[4] = {h.a, h.b, h.c, h.d};
        RayHit r; P3 x;
        if (plane_line_point(hc, p, v, x)) { r.kind = 1; r.pt[0] = x; }
        return r;
    };
    if (orientation(a, b, c, p) != 0 || orientation(a, b, c, q) != 0) {
        const int pqab = orientation(p, q, a, b), pqbc = orientation(p, q, b, c);
        switch (pqab) {
        case 1: return (pqbc != -1 && orientation(p, q, c, a) != -1) ? crossing() : none;
        case -1: return (pqbc != 1 && orientation(p, q, c, a) != 1) ? crossing() : none;
        default:
            if (pqbc == 1) return orientation(p, q, c, a) != -1 ? crossing() : none;
            if (pqbc == -1) return orientation(p, q, c, a) != 1 ? crossing() : none;
            return crossing();
        }
    }
    return line_tri_intersection_coplanar(t, p, q, v);
}
// Plane_3_Triangle_3_intersection.h:27-146: kind 0 none, 1 Point_3, 2 Segment_3, 3 = the Triangle_3 itself (all three vertices on the plane; no points stored)
static inline RayHit plane_tri_intersection(const Tri& t, const double h[4]) {
    const P3* v[3] = {&t.a, &t.b, &t.c};
    const int o0 = plane_oriented_side(h, t.a), o1 = plane_oriented_side(h, t.b), o2 = plane_oriented_side(h, t.c);
    RayHit r;
    auto point = [&](const P3& x) { r.kind = 1; r.pt[0] = x; return r; };
    auto segment = [&](const P3& x, const P3& y) { r.kind = 2; r.pt[0] = x; r.pt[1] = y; return r; };
    auto edge_point = [&](int i, int j, P3& x) { return plane_line_point(h, *v[i], vsub(*v[j], *v[i]), x); };   // Line_3(v_i, v_j): point v_i, direction v_j - v_i
    P3 x;
    if (o0 == 0) {
        if (o1 == 0) { if (o2 == 0) { r.kind = 3; return r; } return segment(t.a, t.b); }
        if (o2 == 0) return segment(t.a, t.c);
        if (o1 == o2) return point(t.a);
        return edge_point(1, 2, x) ? segment(x, t.a) : r;       // (no Point_3: a failed assertion in the reference)
    }
    if (o1 == 0) {
        if (o2 == 0) return segment(t.b, t.c);
        if (o2 == o0) return point(t.b);
        return edge_point(0, 2, x) ? segment(x, t.b) : r;
    }
    if (o2 == 0) {
        if (o1 == o0) return point(t.c);
        return edge_point(0, 1, x) ? segment(x, t.c) : r;
    }
    P3 pts[3]; int np = 0;
    if (o0 != o1) { if (!edge_point(0, 1, x)) return r; pts[np++] = x; }
    if (o0 != o2) { if (!edge_point(0, 2, x)) return r; pts[np++] = x; }
    if (o1 != o2) { if (!edge_point(1, 2, x)) return r; pts[np++] = x; }
    if (np == 0) return r;
    return segment(pts[0], pts[np - 1]);
}

// ------------------------------------------------------------------------------------------------
// Construct_projected_point_3(Triangle_3, Point_3): Kernel_23/include/CGAL/Kernel/function_objects.h:3193-3263
// ------------------------------------------------------------------------------------------------
// Construct_projected_point_3(Line_3, Point_3): Cartesian/function_objects.h:3200-3217 ; line(p1,p2) = (p1, p2-p1) :2690-2691
static inline P3 project_on_line(const P3& lp, const V3& ld, const P3& p) {
    const double dpx = p.x - lp.x, dpy = p.y - lp.y, dpz = p.z - lp.z;
    const double lambda = (ld.x * dpx + ld.y * dpy + ld.z * dpz) / (ld.x * ld.x + ld.y * ld.y + ld.z * ld.z);
    return {lp.x + lambda * ld.x, lp.y + lambda * ld.y, lp.z + lambda * ld.z};
}
// Construct_projected_point_3(Segment_3, Point_3): function_objects.h:3266-3285
static inline P3 project_on_segment(const P3& s, const P3& t, const P3& query) {
    if (s == t) return s;
    const V3 v = vsub(t, s);
    if (dot(v, vsub(query, s)) <= 0) return s;
    if (dot(v, vsub(query, t)) >= 0) return t;
    return project_on_line(s, v, query);
}
static inline P3 project_on_triangle(const Tri& t, const P3& origin) {
    const Plane h = plane_from_points(t.a, t.b, t.c);    // Construct_supporting_plane_3 -> Plane_3(v0,v1,v2)
    if (h.a == 0 && h.b == 0 && h.c == 0) {               // :3208-3247, degenerate triangle -> longest edge by (signed!) Linf
        const V3 ab = vsub(t.b, t.a), ac = vsub(t.c, t.a), bc = vsub(t.c, t.b);
        const double lab = std::max(std::max(ab.x, ab.y), ab.z);
        const double lac = std::max(std::max(ac.x, ac.y), ac.z);
        const double lbc = std::max(std::max(bc.x, bc.y), bc.z);
        if (lab > lac) { if (lab > lbc) return project_on_segment(t.a, t.b, origin); return project_on_segment(t.b, t.c, origin); }
        if (lac > lbc) return project_on_segment(t.a, t.c, origin);
        return project_on_segment(t.b, t.c, origin);
    }
    // projection_planeC3: constructions/kernel_ftC3.h:296-312
    const double num = h.a * origin.x + h.b * origin.y + h.c * origin.z + h.d;
    const double den = h.a * h.a + h.b * h.b + h.c * h.c;
    const double lambda = num / den;
    const P3 proj = {origin.x - lambda * h.a, origin.y - lambda * h.b, origin.z - lambda * h.c};
    const V3 w = {h.a, h.b, h.c};                          // Construct_orthogonal_vector_3(plane)
    // is_inside_triangle_3 :3114-3146 with is_inside_triangle_3_aux :3027-3062 on (t0,t1), (t1,t2), (t2,t0)
    bool outside = false;
    const P3* e[4] = {&t.a, &t.b, &t.c, &t.a};
    for (int i = 0; i < 3; ++i) {
        const P3& p1 = *e[i]; const P3& p2 = *e[i + 1];
        const V3 p1p2 = vsub(p2, p1), p1q = vsub(proj, p1);
        const V3 v = cross(p1p2, p1q);
        if (dot(v, w) < 0) {
            if (dot(p1q, p1p2) >= 0 && dot(vsub(proj, p2), vsub(p1, p2)) >= 0) return project_on_line(p1, p1p2, proj);
            outside = true;
        }
    }
    if (outside) {   // nearest_point_3 :3075-3101
        const double d1 = sqdist(proj, t.a), d2 = sqdist(proj, t.b), d3 = sqdist(proj, t.c);
        if (d2 >= d1 && d3 >= d1) return t.a;
        if (d3 >= d2) return t.b;
        return t.c;
    }
    return proj;
}

// AABB_traits_3::Closest_point: AABB_tree/include/CGAL/AABB_traits_3.h:715-733
static inline P3 closest_point_step(const P3& q, const Tri& t, const P3& bound) {
    const P3 cp = project_on_triangle(t, q);
    return compare_distance(q, cp, bound) == 1 ? bound : cp;
}

// the same for a Segment_3 datum: Construct_projected_point_3(Segment_3, Point_3), function_objects.h:3266-3285
static inline P3 closest_point_step_segment(const P3& q, const P3& a, const P3& b, const P3& bound) {
    const P3 cp = project_on_segment(a, b, q);
    return compare_distance(q, cp, bound) == 1 ? bound : cp;
}

// ------------------------------------------------------------------------------------------------
// Intersection_distance (INEXACT ray/box entry distance): AABB_traits_3.h:89-125
// ------------------------------------------------------------------------------------------------
static inline std::optional<double> intersection_distance(const P3& s, const V3& dir, const Box& bb) {
    double t_near = -DBL_MAX, t_far = DBL_MAX;
    for (int i = 0; i < 3; ++i) {
        if (dir[i] == 0) {
            if (s[i] < bb.lo[i] || s[i] > bb.hi[i]) return std::nullopt;
        } else {
            const double t1 = (bb.lo[i] - s[i]) / dir[i];
            const double t2 = (bb.hi[i] - s[i]) / dir[i];
            t_near = std::max(t_near, std::min(t1, t2));
            t_far = std::min(t_far, std::max(t1, t2));
            if (t_near > t_far || t_far < 0.) return std::nullopt;
        }
    }
    return t_near < 0. ? 0. : t_near;
}

// ------------------------------------------------------------------------------------------------
// Statically_filtered_compare_distance(p, Bbox_3, bound): AABB_traits_3.h:318-459 (generic exact: :131-250)
// returns true iff the traversal must descend (result == SMALLER)
// ------------------------------------------------------------------------------------------------
static inline bool ball_reaches_box_exact(const P3& p, const Box& b, const P3& bound) {
    // do_intersect_sphere_iso_cuboid_3 with exact numbers; sr = |p - bound|^2 exactly
    Expansion sr;
    for (int i = 0; i < 3; ++i) { Expansion d = ex_diff(bound[i], p[i]); sr = sr + d * d; }
    Expansion distance;
    for (int i = 0; i < 3; ++i) {
        Expansion d;
        if (p[i] < b.lo[i]) d = ex_diff(b.lo[i], p[i]);
        else if (p[i] > b.hi[i]) d = ex_diff(p[i], b.hi[i]);
        else continue;
        d = d * d;
        if ((d - sr).sign() > 0) return false;
        distance = distance + d;
    }
    return (distance - sr).sign() <= 0;
}
static inline bool ball_reaches_box(const P3& p, const Box& b, const P3& bound) {
    const double ssr = sqdist(p, bound);
    if (ssr < 1.11261183279326254436e-293 || ssr > 2.80889552322236673473e+306) return ball_reaches_box_exact(p, b, bound);
    double distance = 0, max1 = 0;
    for (int i = 0; i < 3; ++i) {
        double d;
        if (p[i] < b.lo[i]) d = b.lo[i] - p[i];
        else if (p[i] > b.hi[i]) d = p[i] - b.hi[i];
        else continue;
        if (max1 < d) max1 = d;
        distance += d * d;
        const double diff = distance - ssr;
        if (max1 < 3.33558365626356687717e-147 || max1 > 1.67597599124282407923e+153) return true;
        const double eps = 1.99986535548615598560e-15 * std::max(ssr, max1 * max1);
        if (diff > eps) return false;
    }
    return true;
}

}  // namespace orc
