"""Pins the oracle's exact predicates against python rational arithmetic (fractions.Fraction), on random AND
deliberately degenerate inputs (cf. Filtered_kernel/test/Filtered_kernel/Static_filters.cpp:10-13,291-300 and
Intersections_3/test/Intersections_3/bbox_other_do_intersect_test.cpp, which do the same against exact rationals)."""
from fractions import Fraction as Fr

import numpy as np

import orc

SEG, RAY = 0, 1


def F(p):
    return [Fr(float(x)) for x in p]


def sgn(x):
    return (x > 0) - (x < 0)


def det3(a, b, c):
    return (a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]))


def sub(a, b):
    return [x - y for x, y in zip(a, b)]


def orient_fr(p, q, r, s):
    p, q, r, s = F(p), F(q), F(r), F(s)
    return sgn(det3(sub(q, p), sub(r, p), sub(s, p)))


def ulp_nudge(rng, x, k=2):
    for _ in range(int(rng.integers(0, k + 1))):
        x = np.nextafter(x, rng.choice([-np.inf, np.inf]))
    return x


def test_orientation_random_and_degenerate():
    rng = np.random.default_rng(1)
    n_exact_paths = 0
    for it in range(3000):
        mode = it % 4
        if mode == 0:
            pts = rng.uniform(-10, 10, size=(4, 3))
        elif mode == 1:   # integer lattice: many exactly coplanar
            pts = rng.integers(-3, 4, size=(4, 3)).astype(float)
        elif mode == 2:   # s = affine combination of p,q,r (nearly coplanar after rounding)
            pts = rng.uniform(-10, 10, size=(4, 3))
            w = rng.uniform(-1, 2, size=3); w /= w.sum()
            pts[3] = w @ pts[:3]
        else:             # coplanar lattice points nudged by a few ulps, huge offset to stress cancellation
            pts = rng.integers(-3, 4, size=(4, 3)).astype(float) + 1024.0
            pts[3] = np.array([ulp_nudge(rng, x) for x in pts[3]])
        want = orient_fr(*pts)
        assert orc.orientation(*pts) == want
        assert orc.orientation(*pts, exact=True) == want
    assert n_exact_paths == 0


def test_orientation_dir_exact_translation():
    # ray_direction in do_intersect(Ray_3,Triangle_3): sign det[b-a, c-a, q-p] with NO rounding of a+(q-p)
    rng = np.random.default_rng(2)
    for it in range(2000):
        if it % 2:
            a, b, c, p, q = rng.integers(-3, 4, size=(5, 3)).astype(float)
        else:
            a, b, c, p = rng.uniform(-5, 5, size=(4, 3))
            w = rng.uniform(-1, 1, size=2)
            q = p + w[0] * (b - a) + w[1] * (c - a)      # direction nearly in the plane
        fa, fb, fc, fp, fq = F(a), F(b), F(c), F(p), F(q)
        want = sgn(det3(sub(fb, fa), sub(fc, fa), sub(fq, fp)))
        assert orc.orientation_dir(a, b, c, p, q) == want


def test_compare_distance():
    rng = np.random.default_rng(3)
    for it in range(3000):
        p, q = rng.uniform(-10, 10, size=(2, 3))
        if it % 3 == 0:
            r = rng.uniform(-10, 10, size=3)
        elif it % 3 == 1:   # reflection of q through p: equal distance up to rounding
            r = 2 * p - q
        else:
            r = np.array([ulp_nudge(rng, x, 1) for x in q])
        fp, fq, fr = F(p), F(q), F(r)
        want = sgn(sum(x * x for x in sub(fq, fp)) - sum(x * x for x in sub(fr, fp)))
        assert orc.compare_distance(p, q, r) == want
        assert orc.compare_distance(p, q, r, exact=True) == want
    p = np.array([0.5, 0.25, 1.0])
    assert orc.compare_distance(p, np.array([1.5, 0.25, 1.0]), np.array([-0.5, 0.25, 1.0])) == 0


def coplanar_fr(p, q, r):
    p, q, r = F(p), F(q), F(r)
    for i, j in ((0, 1), (1, 2), (0, 2)):
        s = sgn((q[i] - p[i]) * (r[j] - p[j]) - (r[i] - p[i]) * (q[j] - p[j]))
        if s:
            return s
    return 0


def test_coplanar_orientation():
    rng = np.random.default_rng(4)
    for it in range(2000):
        pts = rng.integers(-3, 4, size=(3, 3)).astype(float) if it % 2 else rng.uniform(-1, 1, size=(3, 3))
        if it % 4 == 0:
            pts[2] = pts[0] + rng.uniform(-2, 2) * (pts[1] - pts[0])    # nearly collinear
        assert orc.coplanar_orientation(*pts) == coplanar_fr(*pts)


def seg_box_fr(p, q, box, is_ray):
    """closed segment / ray (p + t(q-p), t in [0,1] or [0,inf)) vs closed box, exact."""
    p, q = F(p), F(q)
    lo, hi = F(box[:3]), F(box[3:])
    t0, t1 = Fr(0), (None if is_ray else Fr(1))
    for i in range(3):
        d = q[i] - p[i]
        if d == 0:
            if p[i] < lo[i] or p[i] > hi[i]:
                return False
            continue
        a, b = (lo[i] - p[i]) / d, (hi[i] - p[i]) / d
        if a > b:
            a, b = b, a
        t0 = max(t0, a)
        t1 = b if t1 is None else min(t1, b)
        if t0 > t1:
            return False
    return True


def test_bbox_segment_ray_do_intersect():
    rng = np.random.default_rng(5)
    for it in range(4000):
        if it % 3 == 0:
            lo = rng.uniform(-1, 0, 3); hi = lo + rng.uniform(0.1, 1, 3)
            pq = rng.uniform(-2, 2, 6)
        elif it % 3 == 1:  # lattice: grazing corners / edges / faces, axis-parallel queries
            lo = rng.integers(-2, 1, 3).astype(float); hi = lo + rng.integers(0, 3, 3)
            pq = rng.integers(-3, 4, 6).astype(float)
        else:              # segment aimed at a box corner, endpoint nudged by ulps
            lo = rng.uniform(-1, 0, 3); hi = lo + rng.uniform(0.1, 1, 3)
            corner = np.where(rng.integers(0, 2, 3) == 1, hi, lo)
            p = rng.uniform(-3, 3, 3)
            q = p + (corner - p) * rng.choice([0.5, 1.0, 2.0])
            pq = np.concatenate([p, [ulp_nudge(rng, x) for x in q]])
        box = np.concatenate([lo, hi])
        for is_ray in (0, 1):
            want = seg_box_fr(pq[:3], pq[3:], box, is_ray)
            if np.all(pq[:3] == pq[3:]):
                continue   # degenerate query: precondition violated in the reference
            assert bool(orc.bbox_do_intersect(box, pq, is_ray)) == want, (box, pq, is_ray)
            assert orc.bbox_do_intersect(box, pq, is_ray, exact=True) == int(want)


def cross_fr(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def dot_fr(a, b):
    return sum(x * y for x, y in zip(a, b))


def line_tri_fr(tri, p, q, is_ray):
    """closed triangle vs closed segment/ray, exact rational arithmetic, independent formulation."""
    a, b, c = F(tri[0:3]), F(tri[3:6]), F(tri[6:9])
    p, q = F(p), F(q)
    n = cross_fr(sub(b, a), sub(c, a))
    assert any(n), "degenerate triangle"
    d = sub(q, p)
    den = dot_fr(n, d)
    num = dot_fr(n, sub(a, p))
    if den != 0:
        t = num / den
        if t < 0 or (not is_ray and t > 1):
            return False
        x = [pi + t * di for pi, di in zip(p, d)]
        # barycentric sign test
        s = [dot_fr(n, cross_fr(sub(v2, v1), sub(x, v1))) for v1, v2 in ((a, b), (b, c), (c, a))]
        return all(v >= 0 for v in s)
    if num != 0:
        return False
    # coplanar: drop the dominant normal axis, 2D test
    k = max(range(3), key=lambda i: abs(n[i]))
    ax = [i for i in range(3) if i != k]
    P2 = lambda v: (v[ax[0]], v[ax[1]])
    A, B, C, P, Q = map(P2, (a, b, c, p, q))

    def o2(u, v, w):
        return sgn((v[0] - u[0]) * (w[1] - u[1]) - (w[0] - u[0]) * (v[1] - u[1]))

    def inside(x):
        s = [o2(A, B, x), o2(B, C, x), o2(C, A, x)]
        return all(v >= 0 for v in s) or all(v <= 0 for v in s)

    if inside(P) or (not is_ray and inside(Q)):
        return True
    D = (Q[0] - P[0], Q[1] - P[1])
    for u, v in ((A, B), (B, C), (C, A)):
        # P + t D = u + s (v-u), t >= 0 (<= 1), s in [0,1]
        e = (v[0] - u[0], v[1] - u[1])
        dn = D[0] * e[1] - D[1] * e[0]
        w = (u[0] - P[0], u[1] - P[1])
        if dn != 0:
            t = (w[0] * e[1] - w[1] * e[0]) / dn
            s = (w[0] * D[1] - w[1] * D[0]) / dn
            if t >= 0 and (is_ray or t <= 1) and 0 <= s <= 1:
                return True
        elif w[0] * D[1] - w[1] * D[0] == 0:   # collinear: project on D
            dd = D[0] * D[0] + D[1] * D[1]
            tu = (w[0] * D[0] + w[1] * D[1]) / dd
            tv = ((v[0] - P[0]) * D[0] + (v[1] - P[1]) * D[1]) / dd
            lo_, hi_ = min(tu, tv), max(tu, tv)
            if hi_ >= 0 and (is_ray or lo_ <= 1):
                return True
    return False


def test_triangle_segment_ray_do_intersect_vs_rationals():
    rng = np.random.default_rng(6)
    n_true = 0
    for it in range(6000):
        if it % 3 == 0:
            tri = rng.uniform(-1, 1, 9); pq = rng.uniform(-1, 1, 6)
        else:   # small lattice: coplanar / touching configurations are frequent
            tri = rng.integers(-2, 3, 9).astype(float); pq = rng.integers(-2, 3, 6).astype(float)
        a, b, c = tri[0:3], tri[3:6], tri[6:9]
        if not np.any(np.cross(b - a, c - a)) or np.all(pq[:3] == pq[3:]):
            continue
        for is_ray in (0, 1):
            want = line_tri_fr(tri, pq[:3], pq[3:], is_ray)
            n_true += want
            assert bool(orc.tri_do_intersect(tri, pq, is_ray)) == want, (tri, pq, is_ray)
            if is_ray:   # the construction-side predicate agrees whenever a + (q-p) is exact (small integers)
                k, pts = orc.tri_ray_intersection(tri, pq)
                if it % 3:
                    assert (k > 0) == want, (tri, pq)
    assert n_true > 500


def test_projection_matches_rational_closest_point():
    # no reference KAT exists for triangle projection; check against the exact closest point, 1e-12 relative
    rng = np.random.default_rng(7)
    for it in range(500):
        tri = rng.uniform(-1, 1, 9)
        q = rng.uniform(-2, 2, 3)
        got = orc.project_on_triangle(tri, q)
        a, b, c = tri[0:3], tri[3:6], tri[6:9]
        # brute force: dense check that no triangle point is closer than `got` (by more than rounding)
        d_got = np.sum((got - q) ** 2)
        u = rng.uniform(0, 1, (2000, 2)); m = u.sum(1) > 1; u[m] = 1 - u[m]
        pts = a + u[:, :1] * (b - a) + u[:, 1:] * (c - a)
        assert d_got <= np.min(np.sum((pts - q) ** 2, 1)) * (1 + 1e-9) + 1e-15
        # and the point lies on the triangle's plane within rounding
        n = np.cross(b - a, c - a)
        assert abs(np.dot(got - a, n)) <= 1e-12 * np.linalg.norm(n) * 4
    # exact KATs: inside -> plane foot, outside edge -> edge foot, outside corner -> vertex
    tri = np.array([0, 0, 0, 1.0, 0, 0, 0, 1.0, 0])
    assert tuple(orc.project_on_triangle(tri, [0.25, 0.25, 7.0])) == (0.25, 0.25, 0.0)
    assert tuple(orc.project_on_triangle(tri, [0.5, -3.0, 2.0])) == (0.5, 0.0, 0.0)
    assert tuple(orc.project_on_triangle(tri, [-1.0, -1.0, 1.0])) == (0.0, 0.0, 0.0)
    assert tuple(orc.project_on_triangle(tri, [3.0, -1.0, 1.0])) == (1.0, 0.0, 0.0)


def test_inexact_box_helpers():
    # Intersection_distance (AABB_traits_3.h:89-125) and the conservative ball/box filter (:318-459)
    box = [0, 0, 0, 1, 1, 1]
    assert orc.intersection_distance([-1, 0.5, 0.5, 0, 0.5, 0.5], box) == 1.0
    assert orc.intersection_distance([0.5, 0.5, 0.5, 1, 0.5, 0.5], box) == 0.0
    assert orc.intersection_distance([-1, 2.0, 0.5, 0, 2.0, 0.5], box) is None
    assert orc.intersection_distance([2, 0.5, 0.5, 3, 0.5, 0.5], box) is None
    assert orc.ball_reaches_box([2, 0.5, 0.5], box, [1.0, 0.5, 0.5]) == 1      # touching: unsure -> descend
    assert orc.ball_reaches_box([2, 0.5, 0.5], box, [1.5, 0.5, 0.5]) == 0
    assert orc.ball_reaches_box([0.5, 0.5, 0.5], box, [0.5, 0.5, 0.5]) == 1
