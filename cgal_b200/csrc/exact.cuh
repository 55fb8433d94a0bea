// exact.cuh — on-device exact sign evaluation (second stage of every filtered predicate).
//
// Replaces the Interval_nt + Exact_nt stages of Filtered_kernel/include/CGAL/Filtered_predicate.h:54-115 by
// floating-point expansion arithmetic (Shewchuk 1997) held in per-thread local arrays of bounded length.
// There is no CPU fallback: a predicate whose static filter fails is resolved here, in the same kernel.
// Exact as long as no intermediate product over/underflows (|coordinates| in ~[1e-140, 1e140]).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_HD_NOINLINE __host__ __device__ __noinline__
#else
#define B2_HD inline
#define B2_HD_NOINLINE
#endif

// The expansion routines are deliberately NOT inlined on the device.  Measured on B200 (scripts/probe/exact_probe3.cu,
// CUDA 12.9): with every routine force-inlined into exact_det3_sign / exact_det2_sign, cicc's optimiser merges the stack
// slots of the per-call scratch arrays incorrectly and the device returns a wrong sign for about half of all inputs
// (host compilation, -G, and any variant with out-of-line exp_add / exp_scale are bit-identical to the host).  Keeping
// them out of line also keeps the rarely executed exact stage out of the traversal kernels' instruction stream.
#if !defined(B2_EXP_LEAF)
#define B2_EXP_LEAF B2_HD_NOINLINE
#endif
#if !defined(B2_EXP_COMP)
#define B2_EXP_COMP B2_HD_NOINLINE
#endif

namespace b2 {

B2_HD void two_sum(double a, double b, double& x, double& y) {
    x = a + b;
    const double bv = x - a;
    const double av = x - bv;
    y = (a - av) + (b - bv);
}
B2_HD void two_diff(double a, double b, double& x, double& y) {
    x = a - b;
    const double bv = a - x;
    const double av = x + bv;
    y = (a - av) + (bv - b);
}
B2_HD void quick_two_sum(double a, double b, double& x, double& y) {  // |a| >= |b|
    x = a + b;
    y = b - (x - a);
}
B2_HD void two_prod(double a, double b, double& x, double& y) {
    x = a * b;
    y = fma(a, b, -x);   // DFMA: exact error term
}

// a value known exactly as hi + lo (result of two_diff)
struct DD { double hi, lo; };
B2_HD DD dd_diff(double a, double b) { DD r; two_diff(a, b, r.hi, r.lo); return r; }

// h = e + f, zero-eliminated; returns length.  h must not alias e or f.
B2_EXP_LEAF int exp_add(const double* e, int ne, const double* f, int nf, double* h) {
    if (ne == 0) { for (int i = 0; i < nf; ++i) h[i] = f[i]; return nf; }
    if (nf == 0) { for (int i = 0; i < ne; ++i) h[i] = e[i]; return ne; }
    int ei = 0, fi = 0, hi = 0;
    double Q, Qn, hh;
    double en = e[0], fn = f[0];
    if ((fn > en) == (fn > -en)) { Q = en; ++ei; } else { Q = fn; ++fi; }
    if (ei < ne && fi < nf) {
        en = e[ei]; fn = f[fi];
        if ((fn > en) == (fn > -en)) { quick_two_sum(en, Q, Qn, hh); ++ei; } else { quick_two_sum(fn, Q, Qn, hh); ++fi; }
        Q = Qn;
        if (hh != 0.0) h[hi++] = hh;
        while (ei < ne && fi < nf) {
            en = e[ei]; fn = f[fi];
            if ((fn > en) == (fn > -en)) { two_sum(Q, en, Qn, hh); ++ei; } else { two_sum(Q, fn, Qn, hh); ++fi; }
            Q = Qn;
            if (hh != 0.0) h[hi++] = hh;
        }
    }
    while (ei < ne) { two_sum(Q, e[ei], Qn, hh); ++ei; Q = Qn; if (hh != 0.0) h[hi++] = hh; }
    while (fi < nf) { two_sum(Q, f[fi], Qn, hh); ++fi; Q = Qn; if (hh != 0.0) h[hi++] = hh; }
    if (Q != 0.0) h[hi++] = Q;
    return hi;
}

// h = e * b, zero-eliminated; returns length (<= 2 ne)
B2_EXP_LEAF int exp_scale(const double* e, int ne, double b, double* h) {
    if (ne == 0 || b == 0.0) return 0;
    int hi = 0;
    double Q, sum, hh, p1, p0;
    two_prod(e[0], b, Q, hh);
    if (hh != 0.0) h[hi++] = hh;
    for (int i = 1; i < ne; ++i) {
        two_prod(e[i], b, p1, p0);
        two_sum(Q, p0, sum, hh);
        if (hh != 0.0) h[hi++] = hh;
        quick_two_sum(p1, sum, Q, hh);
        if (hh != 0.0) h[hi++] = hh;
    }
    if (Q != 0.0) h[hi++] = Q;
    return hi;
}

B2_HD int exp_sign(const double* e, int n) { return n == 0 ? 0 : (e[n - 1] > 0.0 ? 1 : -1); }

// h = a * b for two exact double-doubles; h has room for 8; returns length
B2_EXP_COMP int dd_mul(const DD& a, const DD& b, double* h) {
    double e[2]; int ne = 0;
    if (a.lo != 0.0) e[ne++] = a.lo;
    if (a.hi != 0.0) e[ne++] = a.hi;
    double t1[4], t2[4];
    const int n1 = exp_scale(e, ne, b.lo, t1);
    const int n2 = exp_scale(e, ne, b.hi, t2);
    return exp_add(t1, n1, t2, n2, h);
}

// h = a*b - c*d for exact double-doubles; h has room for 16
B2_EXP_COMP int dd_det2(const DD& a, const DD& b, const DD& c, const DD& d, double* h) {
    double p[8], q[8];
    const int np = dd_mul(a, b, p);
    int nq = dd_mul(c, d, q);
    for (int i = 0; i < nq; ++i) q[i] = -q[i];
    return exp_add(p, np, q, nq, h);
}

// h = e * d for an expansion e (length ne <= 16) and an exact double-double d; h has room for 64; tmp arrays inside
B2_EXP_COMP int exp_mul_dd(const double* e, int ne, const DD& d, double* h) {
    double t1[32], t2[32];
    const int n1 = exp_scale(e, ne, d.lo, t1);
    const int n2 = exp_scale(e, ne, d.hi, t2);
    return exp_add(t1, n1, t2, n2, h);
}

// exact sign of det [a | b | c] for three vectors whose components are exact double-doubles
// (Kernel_23/include/CGAL/determinant.h:41-52 evaluated without rounding)
B2_HD_NOINLINE int exact_det3_sign(DD ax, DD ay, DD az, DD bx, DD by, DD bz, DD cx, DD cy, DD cz) {
    double m[16], t1[64], t2[64], t3[64], s[128], r[192];
    int nm = dd_det2(ax, by, ay, bx, m);            // xy minor
    const int n1 = exp_mul_dd(m, nm, cz, t1);
    nm = dd_det2(ax, bz, az, bx, m);                // xz minor
    int n2 = exp_mul_dd(m, nm, cy, t2);
    for (int i = 0; i < n2; ++i) t2[i] = -t2[i];
    nm = dd_det2(ay, bz, az, by, m);                // yz minor
    const int n3 = exp_mul_dd(m, nm, cx, t3);
    const int ns = exp_add(t1, n1, t2, n2, s);
    const int nr = exp_add(s, ns, t3, n3, r);
    return exp_sign(r, nr);
}

// exact sign of (a*b - c*d) for exact double-doubles
B2_HD_NOINLINE int exact_det2_sign(DD a, DD b, DD c, DD d) {
    double h[16];
    const int n = dd_det2(a, b, c, d, h);
    return exp_sign(h, n);
}

// exact sign of sum_i (q_i - p_i)^2 - (r_i - p_i)^2   (cmp_dist_to_pointC3, kernel_ftC3.h:426-432)
B2_HD_NOINLINE int exact_cmp_dist_sign(const double* p, const double* q, const double* r) {
    double acc[48], tmp[48], sq[8];
    int na = 0;
    for (int i = 0; i < 3; ++i) {
        const DD dq = dd_diff(q[i], p[i]);
        int n = dd_mul(dq, dq, sq);
        int nt = exp_add(acc, na, sq, n, tmp);
        const DD dr = dd_diff(r[i], p[i]);
        n = dd_mul(dr, dr, sq);
        for (int k = 0; k < n; ++k) sq[k] = -sq[k];
        na = exp_add(tmp, nt, sq, n, acc);
    }
    return exp_sign(acc, na);
}

}  // namespace b2
