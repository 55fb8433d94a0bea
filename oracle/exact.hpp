// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under cgal_b200/ may include, link or call this.
//
// Exact sign evaluation for polynomials in IEEE-754 binary64 inputs, by floating-point expansion
// arithmetic (Shewchuk, "Adaptive Precision Floating-Point Arithmetic and Fast Robust Geometric
// Predicates", 1997 — published algorithm, restated here).
//
// Role in the oracle: it stands in for the reference's exact stage, i.e. the third stage of
//   Filtered_kernel/include/CGAL/Filtered_predicate.h:54-115   (static filter -> Interval_nt -> Exact_nt)
// where Exact_nt is Gmpq / boost::multiprecision / Quotient<MP_Float>, all un-vendored externals with no
// pinned version (Number_types/include/CGAL/Number_types/internal/Exact_type_selector.h:141-203).
// The exact stage returns the mathematically exact sign of a polynomial in the input doubles, so any exact
// method yields the same answer; tests/test_oracle_predicates.py pins this file against python `fractions`.
//
// Range: exact as long as no intermediate product over/underflows (|x| roughly in [1e-140, 1e140] for the
// degree-3 predicates).  The reference's rational arithmetic has no such limit; inputs outside the range are
// outside what this oracle certifies ("numeric range").
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

struct Expansion {
    static constexpr int CAP = 400;
    int n = 0;          // number of components, increasing magnitude, non-overlapping, zero-eliminated
    double c[CAP];

    Expansion() {}
    explicit Expansion(double v) { if (v != 0.0) { c[0] = v; n = 1; } }

    int sign() const { return n == 0 ? 0 : (c[n - 1] > 0 ? 1 : -1); }
    double estimate() const { double s = 0; for (int i = 0; i < n; ++i) s += c[i]; return s; }
};

static inline void two_sum(double a, double b, double& x, double& y) {
    x = a + b;
    double bv = x - a;
    double av = x - bv;
    y = (a - av) + (b - bv);
}
static inline void two_diff(double a, double b, double& x, double& y) {
    x = a - b;
    double bv = a - x;
    double av = x + bv;
    y = (a - av) + (bv - b);
}
static inline void fast_two_sum(double a, double b, double& x, double& y) {  // |a| >= |b|
    x = a + b;
    y = b - (x - a);
}
static inline void two_product(double a, double b, double& x, double& y) {
    x = a * b;
    y = std::fma(a, b, -x);  // exact rounding error of the product (IEEE fma)
}

// exact a - b as an expansion of <= 2 components
static inline Expansion ex_diff(double a, double b) {
    Expansion e; double x, y; two_diff(a, b, x, y);
    if (y != 0.0) e.c[e.n++] = y;
    if (x != 0.0) e.c[e.n++] = x;
    return e;
}
static inline Expansion ex_sum2(double a, double b) {
    Expansion e; double x, y; two_sum(a, b, x, y);
    if (y != 0.0) e.c[e.n++] = y;
    if (x != 0.0) e.c[e.n++] = x;
    return e;
}

// h = e + f   (fast_expansion_sum_zeroelim)
static inline void ex_add(const Expansion& e, const Expansion& f, Expansion& h) {
    if (e.n == 0) { h.n = f.n; std::memcpy(h.c, f.c, sizeof(double) * f.n); return; }
    if (f.n == 0) { h.n = e.n; std::memcpy(h.c, e.c, sizeof(double) * e.n); return; }
    double Q, Qnew, hh;
    int ei = 0, fi = 0, hi = 0;
    double enow = e.c[0], fnow = f.c[0];
    if ((fnow > enow) == (fnow > -enow)) { Q = enow; ++ei; } else { Q = fnow; ++fi; }
    if (ei < e.n && fi < f.n) {
        enow = e.c[ei]; fnow = f.c[fi];
        if ((fnow > enow) == (fnow > -enow)) { fast_two_sum(enow, Q, Qnew, hh); ++ei; }
        else { fast_two_sum(fnow, Q, Qnew, hh); ++fi; }
        Q = Qnew;
        if (hh != 0.0) h.c[hi++] = hh;
        while (ei < e.n && fi < f.n) {
            enow = e.c[ei]; fnow = f.c[fi];
            if ((fnow > enow) == (fnow > -enow)) { two_sum(Q, enow, Qnew, hh); ++ei; }
            else { two_sum(Q, fnow, Qnew, hh); ++fi; }
            Q = Qnew;
            if (hh != 0.0) h.c[hi++] = hh;
        }
    }
    while (ei < e.n) { two_sum(Q, e.c[ei], Qnew, hh); ++ei; Q = Qnew; if (hh != 0.0) h.c[hi++] = hh; }
    while (fi < f.n) { two_sum(Q, f.c[fi], Qnew, hh); ++fi; Q = Qnew; if (hh != 0.0) h.c[hi++] = hh; }
    if (Q != 0.0 || hi == 0) { if (Q != 0.0) h.c[hi++] = Q; }
    h.n = hi;
}

// h = e * b   (scale_expansion_zeroelim)
static inline void ex_scale(const Expansion& e, double b, Expansion& h) {
    h.n = 0;
    if (e.n == 0 || b == 0.0) return;
    double Q, sum, hh, p1, p0;
    int hi = 0;
    two_product(e.c[0], b, Q, hh);
    if (hh != 0.0) h.c[hi++] = hh;
    for (int i = 1; i < e.n; ++i) {
        two_product(e.c[i], b, p1, p0);
        two_sum(Q, p0, sum, hh);
        if (hh != 0.0) h.c[hi++] = hh;
        fast_two_sum(p1, sum, Q, hh);
        if (hh != 0.0) h.c[hi++] = hh;
    }
    if (Q != 0.0) h.c[hi++] = Q;
    h.n = hi;
}

static inline Expansion operator+(const Expansion& a, const Expansion& b) { Expansion h; ex_add(a, b, h); return h; }
static inline Expansion operator-(const Expansion& a) { Expansion h; h.n = a.n; for (int i = 0; i < a.n; ++i) h.c[i] = -a.c[i]; return h; }
static inline Expansion operator-(const Expansion& a, const Expansion& b) { return a + (-b); }
static inline Expansion operator*(const Expansion& a, const Expansion& b) {
    Expansion acc, t, u;
    const Expansion& big = (a.n >= b.n) ? a : b;
    const Expansion& small = (a.n >= b.n) ? b : a;
    for (int i = 0; i < small.n; ++i) {
        ex_scale(big, small.c[i], t);
        ex_add(acc, t, u);
        acc = u;
    }
    return acc;
}

}  // namespace orc
