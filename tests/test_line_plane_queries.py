"""Line_3 and Plane_3 intersection queries (the AABB_tree concept's other query types).

1. pins the oracle's restatement (oracle/kernel.hpp: do_intersect_line_tri / _line_bbox / _plane_tri / _plane_bbox) against python rational
   arithmetic, on random and deliberately degenerate inputs — the reference evaluates these Do_intersect_3 overloads exactly
   (Line_3_Triangle_3_do_intersect.h:27-79, Bbox_3_Line_3_do_intersect.h:28-131, Plane_3_Triangle_3_do_intersect.h:27-51,
   Iso_cuboid_3_Plane_3_do_intersect.h:28-125 under Filtered_predicate.h:54-115);
2. runs the DEVICE code (geom.cuh / bvh.cuh compiled for the host) against the oracle: predicate by predicate and through whole traversals.
"""
from fractions import Fraction as Fr

import numpy as np
import pytest

import hc
import orc

LINE, PLANE = 2, 3


def F(p):
    return [Fr(float(x)) for x in p]


def sgn(x):
    return (x > 0) - (x < 0)


def det3(a, b, c):
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])


def sub(a, b):
    return [x - y for x, y in zip(a, b)]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def line_meets_triangle(tri, p, v):
    """ground truth by a DIFFERENT route than the reference's decision tree: the closed triangle and the line p + t v share a point"""
    a, b, c = F(tri[0]), F(tri[1]), F(tri[2])
    p, v = F(p), F(v)
    n = cross(sub(b, a), sub(c, a))
    nv = sum(x * y for x, y in zip(n, v))
    if nv != 0:                                     # one crossing of the supporting plane: is it inside the closed triangle?
        t = sum(x * y for x, y in zip(n, sub(a, p))) / nv
        x = [pi + t * vi for pi, vi in zip(p, v)]
        s = [sgn(sum(i * j for i, j in zip(n, cross(sub(e1, e0), sub(x, e0))))) for e0, e1 in ((a, b), (b, c), (c, a))]
        return all(k >= 0 for k in s)
    if sum(x * y for x, y in zip(n, sub(p, a))) != 0:
        return False                                # parallel and off the plane
    # in the plane: the line separates nothing iff the three vertices are not strictly on one side
    side = [sgn(sum(i * j for i, j in zip(n, cross(v, sub(q, p))))) for q in (a, b, c)]
    return not (all(k > 0 for k in side) or all(k < 0 for k in side))


def line_meets_box(box, p, v):
    lo, hi, p, v = F(box[:3]), F(box[3:]), F(p), F(v)
    t0, t1 = None, None
    for k in range(3):
        if v[k] == 0:
            if p[k] < lo[k] or p[k] > hi[k]:
                return False
            continue
        a, b = (lo[k] - p[k]) / v[k], (hi[k] - p[k]) / v[k]
        a, b = min(a, b), max(a, b)
        t0 = a if t0 is None else max(t0, a)
        t1 = b if t1 is None else min(t1, b)
    return t0 is None or t0 <= t1


def plane_side(h, x):
    h, x = F(h), F(x)
    return sgn(h[0] * x[0] + h[1] * x[1] + h[2] * x[2] + h[3])


def plane_meets_triangle(tri, h):
    s = [plane_side(h, q) for q in tri]
    return not (all(k > 0 for k in s) or all(k < 0 for k in s))


def plane_meets_box(box, h):
    s = [plane_side(h, [box[3 * i], box[3 * j + 1], box[3 * k + 2]]) for i in (0, 1) for j in (0, 1) for k in (0, 1)]
    return not (all(x > 0 for x in s) or all(x < 0 for x in s))


def nudge(rng, x, k=2):
    x = np.array(x, dtype=np.float64)
    for i in range(x.size):
        for _ in range(int(rng.integers(0, k + 1))):
            x.flat[i] = np.nextafter(x.flat[i], rng.choice([-np.inf, np.inf]))
    return x


def line_cases(rng, n):
    """(triangle, point, vector): random, lattice (lines through vertices / along edges / in the plane), and ulp-nudged lattice cases"""
    for it in range(n):
        mode = it % 5
        if mode == 0:
            tri, p, v = rng.uniform(-1, 1, (3, 3)), rng.uniform(-2, 2, 3), rng.uniform(-1, 1, 3)
        elif mode in (1, 2):
            tri = rng.integers(-3, 4, (3, 3)).astype(float)
            while np.linalg.matrix_rank(tri[1:] - tri[0]) < 2:
                tri = rng.integers(-3, 4, (3, 3)).astype(float)
            p = rng.integers(-3, 4, 3).astype(float)
            v = rng.integers(-2, 3, 3).astype(float)
            if mode == 2:                           # through a vertex, along an edge or inside the plane
                k = int(rng.integers(0, 3))
                if k == 0:
                    p = tri[int(rng.integers(0, 3))] - 2 * v
                elif k == 1:
                    v = tri[1] - tri[0]; p = tri[0] + 3 * v
                else:
                    w = rng.integers(-2, 3, 2).astype(float)
                    v = w[0] * (tri[1] - tri[0]) + w[1] * (tri[2] - tri[0]); p = tri[2] + rng.integers(-2, 3) * (tri[1] - tri[0])
        elif mode == 3:                             # the same with a large offset and a few ulps of noise
            off = 1024.0
            tri = rng.integers(-3, 4, (3, 3)).astype(float) + off
            while np.linalg.matrix_rank(tri[1:] - tri[0]) < 2:
                tri = rng.integers(-3, 4, (3, 3)).astype(float) + off
            v = rng.integers(-2, 3, 3).astype(float)
            p = nudge(rng, tri[int(rng.integers(0, 3))] - 2 * v)
        else:                                       # a line aimed exactly at a point of the triangle (rounded): near misses of the edges
            tri = rng.uniform(-1, 1, (3, 3))
            w = rng.dirichlet([0.3, 0.3, 0.3]) if rng.integers(0, 2) else np.array([rng.uniform(), 0, 0])
            if w[1] == 0:
                w[1] = 1 - w[0]
            target = w @ tri
            p = rng.uniform(-2, 2, 3)
            v = target - p
        if not np.any(v):
            v = np.array([1.0, 0.0, 0.0])
        yield tri, p, v


def plane_cases(rng, n):
    for it in range(n):
        mode = it % 4
        if mode == 0:
            tri, h = rng.uniform(-1, 1, (3, 3)), rng.uniform(-1, 1, 4)
        elif mode == 1:                             # lattice: planes through vertices and edges
            tri = rng.integers(-3, 4, (3, 3)).astype(float)
            nrm = rng.integers(-2, 3, 3).astype(float)
            h = np.append(nrm, -nrm @ tri[int(rng.integers(0, 3))])
        elif mode == 2:                             # the plane through a vertex, rounded: sign decided far below the static filter
            tri = rng.uniform(-1, 1, (3, 3)) + 100.0
            nrm = rng.uniform(-1, 1, 3)
            h = np.append(nrm, -(nrm @ tri[int(rng.integers(0, 3))]))
        else:                                       # axis-aligned planes (zero normal components)
            tri = rng.integers(-3, 4, (3, 3)).astype(float)
            nrm = np.zeros(3); nrm[int(rng.integers(0, 3))] = rng.choice([-1.0, 2.0])
            h = np.append(nrm, float(rng.integers(-3, 4)))
        if not np.any(h[:3]):
            h[0] = 1.0
        yield tri, h


def test_oracle_line_triangle_against_rationals():
    rng = np.random.default_rng(11)
    seen = {True: 0, False: 0}
    for tri, p, v in line_cases(rng, 2500):
        want = line_meets_triangle(tri, p, v)
        assert bool(orc.tri_do_intersect_kind(tri, np.concatenate([p, v]), LINE)) == want, (tri, p, v)
        seen[want] += 1
    assert min(seen.values()) > 300


def test_oracle_line_bbox_against_rationals():
    rng = np.random.default_rng(12)
    seen, quirks = {True: 0, False: 0}, 0
    for it in range(3000):
        if it % 3 == 0:
            lo = rng.uniform(-1, 1, 3); hi = lo + rng.uniform(0, 1, 3); p = rng.uniform(-3, 3, 3); v = rng.uniform(-1, 1, 3)
        else:                                       # lattice: lines through corners / edges, axis-parallel directions, flat boxes
            lo = rng.integers(-2, 2, 3).astype(float); hi = lo + rng.integers(0, 3, 3)
            p = rng.integers(-4, 5, 3).astype(float); v = rng.integers(-2, 3, 3).astype(float)
            if it % 3 == 2:
                p = nudge(rng, p + 1024.0) ; lo += 1024.0; hi += 1024.0
        if not np.any(v):
            v[int(rng.integers(0, 3))] = 1.0
        if it % 2:                                  # aim at the box: a corner / face point (lattice) or an interior point
            w = rng.integers(0, 2, 3) if it % 3 else rng.uniform(0, 1, 3)
            p = lo + w * (hi - lo) - float(rng.integers(-3, 4)) * v
            if it % 3 == 2:
                p = nudge(rng, p, 1)
        box = np.concatenate([lo, hi])
        want = line_meets_box(box, p, v)
        got = bool(orc.bbox_do_intersect_kind(box, np.concatenate([p, v]), LINE))
        if got != want:
            # The reference's cross-multiplied slab test (Bbox_3_Line_3_do_intersect.h:53-131) keeps tmax/dmax = 0/0 when a direction
            # component is zero and the point lies exactly on that slab's face; every later comparison then passes.  It is a culling test:
            # a spurious "true" costs a visit and changes no query result.  The oracle restates it as written; never a false "false".
            assert got and not want, (box, p, v)
            assert any(v[k] == 0 and (p[k] == lo[k] or p[k] == hi[k]) for k in range(3)), (box, p, v)
            quirks += 1
        seen[want] += 1
    assert min(seen.values()) > 500 and quirks < 100, (seen, quirks)


def test_oracle_plane_against_rationals():
    rng = np.random.default_rng(13)
    seen = {True: 0, False: 0}
    for tri, h in plane_cases(rng, 3000):
        want = plane_meets_triangle(tri, h)
        assert bool(orc.tri_do_intersect_kind(tri, h, PLANE)) == want, (tri, h)
        seen[want] += 1
        box = np.concatenate([tri.min(0), tri.max(0)])
        assert bool(orc.bbox_do_intersect_kind(box, h, PLANE)) == plane_meets_box(box, h), (box, h)
        far = box + 7.0
        assert bool(orc.bbox_do_intersect_kind(far, h, PLANE)) == plane_meets_box(far, h), (far, h)
    assert min(seen.values()) > 300


def test_device_predicates_against_the_oracle():
    rng = np.random.default_rng(14)
    for tri, p, v in line_cases(rng, 4000):
        q = np.concatenate([p, v])
        assert hc.tri_do_intersect_kind(tri, q, LINE) == orc.tri_do_intersect_kind(tri, q, LINE), (tri, p, v)
    for tri, h in plane_cases(rng, 4000):
        assert hc.tri_do_intersect_kind(tri, h, PLANE) == orc.tri_do_intersect_kind(tri, h, PLANE), (tri, h)
    # magnitudes near the ends of the certified range (DESIGN.md "numeric range")
    for scale in (1e-60, 1e-20, 1e20, 1e60):
        for tri, p, v in line_cases(rng, 300):
            q = np.concatenate([p, v]) * scale
            assert hc.tri_do_intersect_kind(tri * scale, q, LINE) == orc.tri_do_intersect_kind(tri * scale, q, LINE), (scale, tri, p, v)
        for tri, h in plane_cases(rng, 300):
            hs = np.append(h[:3], h[3] * scale)
            assert hc.tri_do_intersect_kind(tri * scale, hs, PLANE) == orc.tri_do_intersect_kind(tri * scale, hs, PLANE), (scale, tri, h)


def _mesh_queries(rng, v, f, n):
    """lines and planes through the mesh: random, through vertices, along edges, in facet planes, axis-parallel"""
    lo, hi = v.min(0), v.max(0)
    ctr, ext = (lo + hi) / 2, (hi - lo).max()
    lines, planes = [], []
    for i in range(n):
        m = i % 5
        if m == 0:
            p = ctr + rng.uniform(-1, 1, 3) * ext; d = rng.normal(size=3)
        elif m == 1:                                # through two mesh vertices
            a, b = v[rng.integers(0, len(v), 2)]
            p, d = a, b - a
        elif m == 2:                                # along an edge, base point far outside
            t = f[rng.integers(0, len(f))]
            d = v[t[1]] - v[t[0]]; p = v[t[0]] - 5 * d
        elif m == 3:                                # axis parallel through a vertex
            p = v[rng.integers(0, len(v))].copy(); d = np.zeros(3); d[rng.integers(0, 3)] = rng.choice([-1.0, 1.0]); p -= 3 * d
        else:                                       # inside a facet's supporting plane (up to rounding)
            t = f[rng.integers(0, len(f))]
            d = rng.uniform(-1, 1) * (v[t[1]] - v[t[0]]) + rng.uniform(-1, 1) * (v[t[2]] - v[t[0]]); p = v[t[0]] + d * 0.25
        if not np.any(d):
            d = np.array([0.0, 0.0, 1.0])
        lines.append(np.concatenate([p, d]))
        if m in (0, 4):
            nrm = rng.normal(size=3); x = ctr + rng.uniform(-0.5, 0.5, 3) * ext
        elif m == 1:
            nrm = rng.normal(size=3); x = v[rng.integers(0, len(v))]
        elif m == 2:                                # the supporting plane of a facet
            t = f[rng.integers(0, len(f))]
            nrm = np.cross(v[t[1]] - v[t[0]], v[t[2]] - v[t[0]]); x = v[t[0]]
        else:
            nrm = np.zeros(3); nrm[rng.integers(0, 3)] = 1.0; x = v[rng.integers(0, len(v))]
        if not np.any(nrm):
            nrm = np.array([1.0, 0.0, 0.0])
        planes.append(np.append(nrm, -(nrm @ x)))
    return np.array(lines), np.array(planes)


@pytest.mark.parametrize("mesh", ["cube", "pinion_small", "lattice"])
def test_device_traversal_counts_against_the_oracle(mesh, meshes):
    rng = np.random.default_rng(15)
    if mesh == "lattice":                           # integer grid of flat triangles: everything coplanar / collinear with everything
        g = np.arange(5, dtype=np.float64)
        v = np.array([[x, y, 0.0] for y in g for x in g] + [[x, 2.0, z] for z in g for x in g])     # two crossing 5 x 5 sheets
        quad = np.array([[j * 5 + i, j * 5 + i + 1, j * 5 + i + 6, j * 5 + i + 5] for j in range(4) for i in range(4)])
        one = np.concatenate([quad[:, [0, 1, 2]], quad[:, [0, 2, 3]]])
        f = np.concatenate([one, one + 25]).astype(np.uint32)
    else:
        v, f = meshes[mesh]
    soup = v[f].reshape(-1, 9)
    lines, planes = _mesh_queries(rng, v, f, 400 if mesh != "pinion_small" else 200)
    o, h = orc.OracleTree(v, f), hc.HostTree(soup)
    for kind, q in ((LINE, lines), (PLANE, planes)):
        want = o.count(q, kind)
        assert np.array_equal(want, o.count(q, kind, naive=True))            # traversal == brute force, inside the oracle
        got = h.count_line_plane(q, kind)
        assert np.array_equal(got, want), (mesh, kind, np.flatnonzero(got != want)[:5])
        assert want.max() > 0
