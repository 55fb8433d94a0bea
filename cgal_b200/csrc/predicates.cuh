// predicates.cuh — Epick predicates as device code: static floating-point filter, then the on-device exact
// stage (exact.cuh).  The filter constants and error models are those of the reference's static filters, so a
// certified sign here is a certified sign there; only the sign (the exact value) is part of the contract.
//   Orientation_3        Filtered_kernel/include/CGAL/Filtered_kernel/internal/Static_filters/Orientation_3.h:158-255
//   Compare_distance_3   .../Static_filters/Compare_distance_3.h:42-134
//   Coplanar_orientation Cartesian_kernel/include/CGAL/predicates/kernel_ftC3.h:196-232 (no static filter in Epick)
#pragma once
#include "exact.cuh"

namespace b2 {

struct Vec3 { double x, y, z; };
B2_HD Vec3 operator-(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
B2_HD Vec3 operator+(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
B2_HD bool same_point(const Vec3& a, const Vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
B2_HD double comp(const Vec3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
B2_HD double fmax3(double a, double b, double c) { return fmax(a, fmax(b, c)); }
B2_HD double fmin3(double a, double b, double c) { return fmin(a, fmin(b, c)); }

// sign of det[u | v | w] where u,v,w are (rounded) differences of input coordinates and (eu,ev,ew) recompute
// them exactly on demand.  Semi-static filter of Orientation_3.h:219-251.
template <class ExactFn>
B2_HD int filtered_det3(const Vec3& u, const Vec3& v, const Vec3& w, int& nexact, ExactFn exact) {
    const double maxx = fmax3(fabs(u.x), fabs(v.x), fabs(w.x));
    const double maxy = fmax3(fabs(u.y), fabs(v.y), fabs(w.y));
    const double maxz = fmax3(fabs(u.z), fabs(v.z), fabs(w.z));
    const double m01 = u.x * v.y - v.x * u.y;
    const double m02 = u.x * w.y - w.x * u.y;
    const double m12 = v.x * w.y - w.x * v.y;
    const double det = m01 * w.z - m02 * v.z + m12 * u.z;
    const double eps = 5.1107127829973299e-15 * maxx * maxy * maxz;
    const double lo = fmin3(maxx, maxy, maxz), hi = fmax3(maxx, maxy, maxz);
    if (lo >= 1e-97) {
        if (hi < 1e102) {
            if (det > eps) return 1;
            if (det < -eps) return -1;
        }
    } else if (lo == 0.0) {
        return 0;   // a zero row: the determinant is exactly zero
    }
    ++nexact;
    return exact();
}

// Orientation_3(p,q,r,s) = sign det[q-p | r-p | s-p]
B2_HD int orient3(const Vec3& p, const Vec3& q, const Vec3& r, const Vec3& s, int& nexact) {
    return filtered_det3(q - p, r - p, s - p, nexact, [&]() {
        return exact_det3_sign(dd_diff(q.x, p.x), dd_diff(q.y, p.y), dd_diff(q.z, p.z),
                               dd_diff(r.x, p.x), dd_diff(r.y, p.y), dd_diff(r.z, p.z),
                               dd_diff(s.x, p.x), dd_diff(s.y, p.y), dd_diff(s.z, p.z));
    });
}

// sign det[b-a | c-a | q-p]: Orientation_3(a,b,c, a+(q-p)) with the translation carried out exactly, which is what
// the reference evaluates for `ray_direction` inside do_intersect(Ray_3,Triangle_3)
// (Intersections_3/internal/Ray_3_Triangle_3_do_intersect.h:218-219 under Filtered_predicate)
B2_HD int orient3_dir(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& p, const Vec3& q, int& nexact) {
    return filtered_det3(b - a, c - a, q - p, nexact, [&]() {
        return exact_det3_sign(dd_diff(b.x, a.x), dd_diff(b.y, a.y), dd_diff(b.z, a.z),
                               dd_diff(c.x, a.x), dd_diff(c.y, a.y), dd_diff(c.z, a.z),
                               dd_diff(q.x, p.x), dd_diff(q.y, p.y), dd_diff(q.z, p.z));
    });
}

// sign det[v | a-p | b-p] = Orientation_3(p, p + v, a, b) with the point p + v formed EXACTLY: what Filtered_predicate evaluates for the
// Line_3 queries (a Line_3 is stored as point + direction vector, Cartesian/Line_3.h; Line_3_Triangle_3_do_intersect.h:45-46 takes
// point(0) and point(1) = p + v inside the exact kernel)
B2_HD int orient3_vec(const Vec3& v, const Vec3& p, const Vec3& a, const Vec3& b, int& nexact) {
    return filtered_det3(v, a - p, b - p, nexact, [&]() {
        return exact_det3_sign(DD{v.x, 0.0}, DD{v.y, 0.0}, DD{v.z, 0.0},
                               dd_diff(a.x, p.x), dd_diff(a.y, p.y), dd_diff(a.z, p.z),
                               dd_diff(b.x, p.x), dd_diff(b.y, p.y), dd_diff(b.z, p.z));
    });
}

// orientationC2 (kernel_ftC2.h:403-408): sign((qx-px)(ry-py) - (rx-px)(qy-py))
B2_HD int orient2(double px, double py, double qx, double qy, double rx, double ry, int& nexact) {
    const double a = qx - px, b = qy - py, c = rx - px, d = ry - py;
    const double l = a * d, r = c * b, det = l - r;
    const double mag = fmax(fmax(fabs(a), fabs(b)), fmax(fabs(c), fabs(d)));
    if (mag > 1e-140 && mag < 1e140) {
        const double err = 8.8817841970012523e-16 * (fabs(l) + fabs(r));
        if (det > err) return 1;
        if (det < -err) return -1;
    }
    ++nexact;
    return exact_det2_sign(dd_diff(qx, px), dd_diff(ry, py), dd_diff(rx, px), dd_diff(qy, py));
}
// coplanar_orientationC3, 3-point version (kernel_ftC3.h:218-232)
B2_HD int coplanar_orient(const Vec3& p, const Vec3& q, const Vec3& r, int& nexact) {
    int o = orient2(p.x, p.y, q.x, q.y, r.x, r.y, nexact);
    if (o != 0) return o;
    o = orient2(p.y, p.z, q.y, q.z, r.y, r.z, nexact);
    if (o != 0) return o;
    return orient2(p.x, p.z, q.x, q.z, r.x, r.z, nexact);
}
// the same with q = p + v formed exactly (Line_3 queries): sign(vx (ry - py) - (rx - px) vy), and the coplanar_orientationC3 cascade over it
B2_HD int orient2_vec(double vx, double vy, double px, double py, double rx, double ry, int& nexact) {
    const double c = rx - px, d = ry - py;
    const double l = vx * d, r = c * vy, det = l - r;
    const double mag = fmax(fmax(fabs(vx), fabs(vy)), fmax(fabs(c), fabs(d)));
    if (mag > 1e-140 && mag < 1e140) {
        const double err = 8.8817841970012523e-16 * (fabs(l) + fabs(r));
        if (det > err) return 1;
        if (det < -err) return -1;
    }
    ++nexact;
    return exact_det2_sign(DD{vx, 0.0}, dd_diff(ry, py), dd_diff(rx, px), DD{vy, 0.0});
}
B2_HD int coplanar_orient_vec(const Vec3& p, const Vec3& v, const Vec3& r, int& nexact) {
    int o = orient2_vec(v.x, v.y, p.x, p.y, r.x, r.y, nexact);
    if (o != 0) return o;
    o = orient2_vec(v.y, v.z, p.y, p.z, r.y, r.z, nexact);
    if (o != 0) return o;
    return orient2_vec(v.x, v.z, p.x, p.z, r.x, r.z, nexact);
}
// Oriented_side_3(Plane_3, Point_3) = sign(a x + b y + c z + d) (side_of_oriented_planeC3, kernel_ftC3.h), exact: a filter on the rounded sum,
// then the seven-component expansion of the three products and d
B2_HD_NOINLINE int plane_side_exact(double a, double b, double c, double d, double x, double y, double z) {
    double t1[2], t2[2], t3[2], s1[4], s2[6], s3[8];
    int n1 = 0, n2 = 0, n3 = 0;
    double hi, lo;
    two_prod(a, x, hi, lo); if (lo != 0.0) t1[n1++] = lo; if (hi != 0.0) t1[n1++] = hi;
    two_prod(b, y, hi, lo); if (lo != 0.0) t2[n2++] = lo; if (hi != 0.0) t2[n2++] = hi;
    two_prod(c, z, hi, lo); if (lo != 0.0) t3[n3++] = lo; if (hi != 0.0) t3[n3++] = hi;
    const int m1 = exp_add(t1, n1, t2, n2, s1);
    const int m2 = exp_add(s1, m1, t3, n3, s2);
    double dd[1] = {d};
    const int m3 = exp_add(s2, m2, dd, d != 0.0 ? 1 : 0, s3);
    return exp_sign(s3, m3);
}
B2_HD int plane_side(double a, double b, double c, double d, const Vec3& p, int& nexact) {
    const double ax = a * p.x, by = b * p.y, cz = c * p.z;
    const double v = ((ax + by) + cz) + d;
    const double mag = ((fabs(ax) + fabs(by)) + fabs(cz)) + fabs(d);
    if (mag < 1e290 && (mag > 1e-290 || mag == 0.0)) {
        const double err = 6.7e-16 * mag;                     // three products + three sums: < 6 * 2^-53 relative to the magnitude sum
        if (v > err) return 1;
        if (v < -err) return -1;
        if (mag == 0.0) return 0;
    }
    ++nexact;
    return plane_side_exact(a, b, c, d, p.x, p.y, p.z);
}

// collinear_are_ordered_along_lineC3 (kernel_ftC3.h:279-291)
B2_HD bool collinear_ordered(const Vec3& p, const Vec3& q, const Vec3& r) {
    if (p.x < q.x) return !(r.x < q.x);
    if (q.x < p.x) return !(q.x < r.x);
    if (p.y < q.y) return !(r.y < q.y);
    if (q.y < p.y) return !(q.y < r.y);
    if (p.z < q.z) return !(r.z < q.z);
    if (q.z < p.z) return !(q.z < r.z);
    return true;
}

// Second filter stage of Compare_distance_3 (takes the place of the reference's interval stage, Filtered_predicate.h:54-115):
//   |q-p|^2 - |r-p|^2 = sum_i (q_i - r_i) (q_i + r_i - 2 p_i)
// evaluated in double-double.  e_i = q_i - r_i is an exact pair (two_diff); s_i = q_i + r_i - 2 p_i is an exact triple
// (two_sum, two_diff) folded into a pair with relative error 2^-105; each product and the 3-term sum add < 2^-100 relative
// to sum_i |e_i s_i|, so a result larger than 2^-90 times that sum has a certain sign.  Returns 2 when unsure (exact ties
// between different points, or magnitudes near the under/overflow thresholds): the caller then runs the expansion stage.
// Typical use: candidates of two triangles on a shared edge differ by a few ulps, which defeats the 3.8e-15 static filter
// for about a third of all closest_point queries on a mesh but is 2^50 above this stage's resolution.
B2_HD int cmp_dist_dd(const double* p, const double* q, const double* r) {
    double hi = 0.0, lo = 0.0, mag = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double eh, el, sh, sl, th, tl;
        two_diff(q[i], r[i], eh, el);
        two_sum(q[i], r[i], sh, sl);
        two_diff(sh, 2.0 * p[i], th, tl);
        tl = tl + sl;
        double ph, pl;
        two_prod(eh, th, ph, pl);
        pl = pl + (eh * tl + el * th);
        mag += fabs(ph);
        double nh, nl;
        two_sum(hi, ph, nh, nl);
        hi = nh; lo = lo + (nl + pl);
    }
    const double d = hi + lo;
    if (!(mag > 1e-250 && mag < 1e250)) return 2;
    const double err = mag * 8.1e-28;   // 2^-90
    if (d > err) return 1;
    if (d < -err) return -1;
    return 2;
}

// Compare_distance_3(p,q,r) = sign(|q-p|^2 - |r-p|^2)
B2_HD int cmp_dist(const Vec3& p, const Vec3& q, const Vec3& r, int& nexact) {
    if (same_point(q, r)) return 0;
    const Vec3 dq = q - p, dr = r - p;
    const double qp2 = (dq.x * dq.x + dq.y * dq.y) + dq.z * dq.z;
    const double rp2 = (dr.x * dr.x + dr.y * dr.y) + dr.z * dr.z;
    const double max1 = fmax(fmax3(fabs(dq.x), fabs(dq.y), fabs(dq.z)), fmax3(fabs(dr.x), fabs(dr.y), fabs(dr.z)));
    if (max1 >= 2.42701102401884262260e-147 && max1 <= 8.37987995621411946582e+152) {
        const double eps = 3.77746921267322435884e-15 * (max1 * max1);
        const double diff = qp2 - rp2;
        if (diff > eps) return 1;
        if (diff < -eps) return -1;
    }
    const double pa[3] = {p.x, p.y, p.z}, qa[3] = {q.x, q.y, q.z}, ra[3] = {r.x, r.y, r.z};
    const int s2 = cmp_dist_dd(pa, qa, ra);
    if (s2 != 2) return s2;
    ++nexact;
    return exact_cmp_dist_sign(pa, qa, ra);
}

}  // namespace b2
