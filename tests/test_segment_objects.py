"""Intersect_3(Segment_3, Triangle_3) (Intersections_3/include/CGAL/Intersections_3/internal/Segment_3_Triangle_3_intersection.h): the intersection OBJECTS
of Segment_3 queries (AABB_tree::all_intersections / any_intersection, AABB_tree.h:958-985).
1. the oracle's restatement against an independent clipping in rational arithmetic (which object, and where, to 1e-12 of the coordinates' magnitude);
2. the device routine (geom.cuh::seg_tri_intersection, compiled for the host) against the oracle, bit for bit."""
from fractions import Fraction as Fr

import numpy as np

import hc
import orc


def F(p):
    return [Fr(float(x)) for x in p]


def sub(a, b):
    return [x - y for x, y in zip(a, b)]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def exact_intersection(tri, seg):
    """(kind, [points]) of the closed segment with the closed triangle, in rationals: Cyrus-Beck clipping of the parameter interval"""
    a, b, c = F(tri[0:3]), F(tri[3:6]), F(tri[6:9])
    p, q = F(seg[0:3]), F(seg[3:6])
    d = sub(q, p)
    n = cross(sub(b, a), sub(c, a))
    den, num = dot(n, d), dot(n, sub(a, p))
    at = lambda t: [pi + t * di for pi, di in zip(p, d)]     # noqa: E731
    inside = lambda x: all(dot(n, cross(sub(e1, e0), sub(x, e0))) >= 0 for e0, e1 in ((a, b), (b, c), (c, a)))   # noqa: E731
    if den != 0:
        t = num / den
        if t < 0 or t > 1 or not inside(at(t)):
            return 0, []
        return 1, [at(t)]
    if num != 0:
        return 0, []
    t0, t1 = Fr(0), Fr(1)                                    # coplanar: clip against the three edges
    for e0, e1 in ((a, b), (b, c), (c, a)):
        f0 = dot(n, cross(sub(e1, e0), sub(p, e0)))          # >= 0 inside
        f1 = dot(n, cross(sub(e1, e0), sub(q, e0)))
        if f0 < 0 and f1 < 0:
            return 0, []
        if f0 < 0:
            t0 = max(t0, f0 / (f0 - f1))
        elif f1 < 0:
            t1 = min(t1, f0 / (f0 - f1))
    if t0 > t1:
        return 0, []
    if t0 == t1:
        return 1, [at(t0)]
    return 2, [at(t0), at(t1)]


def lattice(rng, n, lo=-3, hi=4):
    return rng.integers(lo, hi, n).astype(np.float64)


def cases(rng, n):
    for it in range(n):
        if it % 4 == 3:   # everything in one axis-parallel or oblique lattice plane: the coplanar case analyses
            tri, pq = lattice(rng, 9), lattice(rng, 6)
            ax = it % 3
            tri[ax::3] = 1.0; pq[ax::3] = 1.0
            if it % 8 == 7:
                tri[2::3] = 3 - tri[0::3] - tri[1::3]; pq[2::3] = 3 - pq[0::3] - pq[1::3]
            if it % 16 == 15:                                # collinear with an edge of the triangle
                k = rng.integers(-2, 3, 2)
                pq = np.concatenate([tri[0:3] + k[0] * (tri[3:6] - tri[0:3]), tri[0:3] + k[1] * (tri[3:6] - tri[0:3])])
        elif it % 3 == 0:
            tri, pq = rng.uniform(-1, 1, 9), rng.uniform(-1, 1, 6)
        elif it % 3 == 1:
            tri, pq = lattice(rng, 9), lattice(rng, 6)
        else:             # a segment through (or ending at) a point of the triangle
            tri = rng.uniform(-1, 1, 9)
            target = rng.dirichlet([1, 1, 1]) @ tri.reshape(3, 3)
            s = rng.uniform(-3, 3, 3)
            pq = np.concatenate([s, s + (target - s) * rng.choice([0.5, 1.0, 1.0, 2.0, 3.0])])
        a, b, c = tri[:3], tri[3:6], tri[6:]
        if not np.any(np.cross(b - a, c - a)) or np.all(pq[:3] == pq[3:]):
            continue
        yield tri, pq


def test_oracle_against_rational_clipping():
    rng = np.random.default_rng(51)
    seen = np.zeros(3, int)
    for tri, pq in cases(rng, 6000):
        k, pts = orc.tri_seg_intersection(tri, pq)
        assert (k > 0) == bool(orc.tri_do_intersect(tri, pq, 0)), (tri, pq)      # the object exists exactly when the predicate says so
        exact_in = not np.any(tri != np.round(tri)) and not np.any(pq != np.round(pq))
        if not exact_in:
            continue                                                              # rounded inputs: the predicate decides, positions checked on lattice input
        ke, pe = exact_intersection(tri, pq)
        assert k == ke, (tri, pq, k, ke)
        seen[k] += 1
        got = sorted(tuple(x) for x in pts[:k])
        want = sorted(tuple(float(v) for v in x) for x in pe)
        np.testing.assert_allclose(np.array(got).reshape(-1), np.array(want).reshape(-1), rtol=0, atol=1e-12)
    assert seen.min() > 100, seen


def test_device_routine_against_the_oracle():
    rng = np.random.default_rng(52)
    seen = np.zeros(3, int)
    for tri, pq in cases(rng, 12000):
        k, pts = hc.tri_seg_intersection(tri, pq)
        ko, ptso = orc.tri_seg_intersection(tri, pq)
        assert k == ko, (tri, pq)
        np.testing.assert_array_equal(pts[:k], ptso[:k])                          # bit-exact, endpoint order included
        seen[k] += 1
    assert seen.min() > 500, seen


# ---- Line_3 and Plane_3 queries: Line_3_Triangle_3_intersection.h:27-419, Plane_3_Triangle_3_intersection.h:27-146 ----------------------------------
def exact_line_intersection(tri, p, v):
    """the closed triangle with the line p + t v, in rationals (p + v taken exactly: on lattice input the rounded second point IS exact)"""
    k, pts = exact_intersection(tri, np.concatenate([p - 1000 * v, p + 1000 * v]))      # lattice triangles live in [-3, 3]^3: a segment long enough
    return k, pts


def exact_plane_intersection(tri, h):
    """the closed triangle with the plane: vertices on it and the crossing points of the edges that change side"""
    V = [F(tri[0:3]), F(tri[3:6]), F(tri[6:9])]
    H = F(h)
    val = [H[0] * x[0] + H[1] * x[1] + H[2] * x[2] + H[3] for x in V]
    on = [i for i in range(3) if val[i] == 0]
    if len(on) == 3:
        return 3, []
    pts = [V[i] for i in on]
    for i, j in ((0, 1), (0, 2), (1, 2)):
        if val[i] * val[j] < 0:
            t = val[i] / (val[i] - val[j])
            pts.append([a + t * (b - a) for a, b in zip(V[i], V[j])])
    if not pts:
        return 0, []
    return (1, pts) if len(pts) == 1 else (2, pts)


def test_line_and_plane_objects_against_rationals_and_device_against_oracle():
    from test_line_plane_queries import line_cases, plane_cases
    rng = np.random.default_rng(53)
    seen_l, seen_p = np.zeros(3, int), np.zeros(4, int)
    for tri, p, v in line_cases(rng, 6000):
        tri = tri.reshape(9)
        q = np.concatenate([p, v])
        k, pts = orc.tri_query_intersection(tri, q, 2)
        kd, ptsd = hc.tri_query_intersection(tri, q, 2)
        assert kd == k, (tri, p, v)
        np.testing.assert_array_equal(ptsd[:k], pts[:k])                                 # device = oracle, bit for bit
        if np.any(tri != np.round(tri)) or np.any(q != np.round(q)):
            continue
        ke, pe = exact_line_intersection(tri, p, v)
        assert k == ke, (tri, p, v, k, ke)
        seen_l[k] += 1
        got = sorted(tuple(x) for x in pts[:k])
        want = sorted(tuple(float(c) for c in x) for x in pe)
        np.testing.assert_allclose(np.array(got).reshape(-1), np.array(want).reshape(-1), rtol=0, atol=1e-12)
    for tri, h in plane_cases(rng, 6000):
        tri = tri.reshape(9)
        k, pts = orc.tri_query_intersection(tri, h, 3)
        kd, ptsd = hc.tri_query_intersection(tri, h, 3)
        assert kd == k, (tri, h)
        if k in (1, 2):
            np.testing.assert_array_equal(ptsd[:k], pts[:k])
        assert (k > 0) == bool(orc.tri_do_intersect_kind(tri, h, 3)), (tri, h)           # an object exactly when the predicate says so
        if np.any(tri != np.round(tri)) or np.any(h != np.round(h)):
            continue
        ke, pe = exact_plane_intersection(tri, h)
        assert k == ke, (tri, h, k, ke)
        seen_p[k] += 1
        if k in (1, 2):
            got = sorted(tuple(x) for x in pts[:k])
            want = sorted(tuple(float(c) for c in x) for x in pe)
            np.testing.assert_allclose(np.array(got).reshape(-1), np.array(want).reshape(-1), rtol=0, atol=1e-12)
    assert seen_l.min() > 50 and seen_p[:3].min() > 50, (seen_l, seen_p)
