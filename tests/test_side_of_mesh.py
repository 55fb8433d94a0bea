"""Side_of_triangle_mesh (Polygon_mesh_processing/include/CGAL/Side_of_triangle_mesh.h:265-306): point in / on / outside a
closed triangle mesh by ray casting with exact predicates and re-tries on degenerate rays.

1. the oracle's restatement (vertical-ray traits + generic-ray traits + retries) against an independent classification in rational arithmetic;
2. the device code (geom.cuh::ray_tri_classify, bvh.cuh::traverse_side, compiled for the host) against the oracle — with a DIFFERENT sequence of
   rays, because the answer must not depend on them."""
from fractions import Fraction as Fr

import numpy as np
import pytest

import hc
import orc

DIRS = np.array([[0, 0, 1.0], [0, 0, -1.0], [0.3, -0.5, 0.8], [-0.7, 0.2, 0.1], [0.123, 0.456, -0.789], [0.9, 0.31, 0.17], [-0.21, -0.83, 0.4], [0.05, 0.6, 0.33]])


def F(p):
    return [Fr(float(x)) for x in p]


def sub(a, b):
    return [x - y for x, y in zip(a, b)]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def on_triangle(p, a, b, c):
    n = cross(sub(b, a), sub(c, a))
    if dot(n, sub(p, a)) != 0:
        return False
    s = [dot(n, cross(sub(e1, e0), sub(p, e0))) for e0, e1 in ((a, b), (b, c), (c, a))]
    return all(x >= 0 for x in s)


def ray_parity(p, d, tris):
    """number of facets the open ray p + t d (t > 0) crosses in their interior, or None when it touches an edge / a vertex or runs inside a facet's plane"""
    count = 0
    for a, b, c in tris:
        n = cross(sub(b, a), sub(c, a))
        den = dot(n, d)
        num = dot(n, sub(a, p))
        if den == 0:
            if num != 0:
                continue
            # the ray runs in the facet's plane: degenerate iff it meets the closed triangle; decide conservatively (any vertex not strictly on one side)
            side = [dot(n, cross(d, sub(v, p))) for v in (a, b, c)]
            if all(s > 0 for s in side) or all(s < 0 for s in side):
                continue
            return None
        t = num / den
        if t <= 0:
            continue
        x = [pi + t * di for pi, di in zip(p, d)]
        s = [dot(n, cross(sub(e1, e0), sub(x, e0))) for e0, e1 in ((a, b), (b, c), (c, a))]
        if any(v < 0 for v in s):
            continue
        if any(v == 0 for v in s):
            return None
        count += 1
    return count


def side_truth(pt, soup, lo, hi):
    """-1 / 0 / +1 in rational arithmetic; a float prefilter only skips facets whose box misses an axis-parallel ray"""
    if np.any(pt < lo) or np.any(pt > hi):
        inside_box = False
    else:
        inside_box = True
    p = F(pt)
    near = np.all(soup.reshape(-1, 3, 3).min(1) <= pt, axis=1) & np.all(soup.reshape(-1, 3, 3).max(1) >= pt, axis=1)
    for t in soup[near]:
        if on_triangle(p, F(t[0:3]), F(t[3:6]), F(t[6:9])):
            return 0
    if not inside_box:
        return -1
    T = soup.reshape(-1, 3, 3)
    for axis in (2, 0, 1):
        others = [k for k in range(3) if k != axis]
        m = np.ones(len(T), bool)
        for k in others:
            m &= (T[:, :, k].min(1) <= pt[k]) & (T[:, :, k].max(1) >= pt[k])
        for sgn in (1, -1):
            d = [Fr(0)] * 3
            d[axis] = Fr(sgn)
            c = ray_parity(p, d, [(F(t[0]), F(t[1]), F(t[2])) for t in T[m]])
            if c is not None:
                return 1 if c % 2 else -1
    for d in ([Fr(3, 7), Fr(-5, 11), Fr(8, 13)], [Fr(-7, 9), Fr(2, 17), Fr(1, 19)], [Fr(1, 3), Fr(1, 5), Fr(-1, 7)]):
        c = ray_parity(p, d, [(F(t[0]), F(t[1]), F(t[2])) for t in T])
        if c is not None:
            return 1 if c % 2 else -1
    raise AssertionError("no generic ray found")


def special_points(rng, v, f, n_random):
    """random points in a slightly enlarged box + points ON the mesh and points whose vertical ray is degenerate"""
    lo, hi = v.min(0), v.max(0)
    ext = hi - lo
    pts = [lo - 0.1 * ext + rng.random((n_random, 3)) * 1.2 * ext]
    k = rng.integers(0, len(f), 60)
    tri = v[f[k].astype(np.int64)]
    pts.append(v[rng.integers(0, len(v), 40)])                                  # vertices: on the boundary
    pts.append(0.5 * (tri[:, 0] + tri[:, 1]))                                   # edge midpoints (on the boundary when representable)
    pts.append(tri.mean(1))                                                     # facet centroids (rounded: anywhere near the facet)
    above = v[rng.integers(0, len(v), 40)].copy()                               # directly above / below a vertex: the vertical ray hits it
    above[:, 2] += rng.uniform(-0.3, 0.3, len(above)) * ext[2]
    pts.append(above)
    mid = 0.5 * (tri[:, 1] + tri[:, 2])                                         # above / below an edge midpoint
    mid[:, 2] += rng.uniform(-0.3, 0.3, len(mid)) * ext[2]
    pts.append(mid)
    return np.concatenate(pts)


def lattice_cube():
    """a 2 x 2 x 2 cube of 48 triangles with integer coordinates: facet planes, edges and vertices are hit exactly by lattice points and axis rays"""
    from itertools import product
    quads = []
    for axis in range(3):
        o = [k for k in range(3) if k != axis]
        for side in (0, 2):
            for i, j in product((0, 1), (0, 1)):
                q = []
                for di, dj in ((0, 0), (1, 0), (1, 1), (0, 1)):
                    p = [0, 0, 0]
                    p[axis] = side; p[o[0]] = i + di; p[o[1]] = j + dj
                    q.append(p)
                quads.append(q if side == 2 else q[::-1])
    soup = []
    for q in quads:
        soup += [q[0] + q[1] + q[2], q[0] + q[2] + q[3]]
    return np.array(soup, dtype=np.float64)


def test_oracle_against_rationals(meshes):
    rng = np.random.default_rng(31)
    seen = {-1: 0, 0: 0, 1: 0}
    cases = [("cube", None), ("tetrahedron", None), ("lattice", lattice_cube()), ("pinion_small", None)]
    for name, soup in cases:
        if soup is None:
            v, f = meshes[name]
            soup = v[f.astype(np.int64)].reshape(-1, 9)
        else:
            v, f = soup.reshape(-1, 3), np.arange(len(soup) * 3).reshape(-1, 3)
        o = orc.OracleTree(soup=soup)
        if name == "lattice":
            g = np.arange(-1, 7) * 0.5
            pts = np.array([[x, y, z] for x in g for y in g for z in g])             # every lattice point: corners, edges, faces, interior
        else:
            pts = special_points(rng, v, f, 150 if name != "pinion_small" else 40)[: (10 ** 9 if name != "pinion_small" else 120)]
        side, attempts = o.side_of_mesh(pts)
        lo, hi = soup.reshape(-1, 3).min(0), soup.reshape(-1, 3).max(0)
        for p, s in zip(pts, side):
            want = side_truth(p, soup, lo, hi)
            assert s == want, (name, p, s, want)
            seen[want] += 1
        if name == "lattice":
            assert attempts.max() > 1                                                  # lattice points force re-tries
    assert min(seen.values()) > 40, seen


def test_device_ray_triangle_classification_against_the_oracle():
    from test_line_plane_queries import line_cases
    rng = np.random.default_rng(32)
    kinds = np.zeros(6, int)
    for tri, p, v in line_cases(rng, 6000):
        q = np.concatenate([p, p + v])
        want = orc.classify_ray_tri(tri, q)
        assert hc.classify_ray_tri(tri, q) == want, (tri, p, v)
        kinds[want + 1] += 1
    assert kinds.min() > 20, kinds                                                     # none / facet / edge / vertex / coplanar ray / endpoint all seen


@pytest.mark.parametrize("name", ["cube", "tetrahedron", "lattice", "elephant", "handle", "coverrear"])
def test_device_traversal_against_the_oracle(name, meshes):
    rng = np.random.default_rng(33)
    if name == "lattice":
        soup = lattice_cube()
        g = np.arange(-1, 7) * 0.5
        pts = np.array([[x, y, z] for x in g for y in g for z in g])
    else:
        v, f = meshes[name]
        soup = v[f.astype(np.int64)].reshape(-1, 9)
        pts = special_points(rng, v, f, 3000 if len(f) > 100 else 500)
    want, _ = orc.OracleTree(soup=soup).side_of_mesh(pts)
    got, attempts = hc.HostTree(soup).side_of_mesh(pts, DIRS)
    assert np.array_equal(got, want), (name, np.flatnonzero(got != want)[:5])
    assert (want == 1).any() and (want == -1).any()
    if name in ("lattice", "cube"):
        assert (want == 0).any() and attempts.max() > 1


def transformed_soup(m, soup):
    """images of the primitives as Aff_transformation_3::transform gives them (Cartesian/Aff_transformation_rep_3.h:111-117): left to right in binary64"""
    p = soup.reshape(-1, 3)
    m = np.asarray(m, float).reshape(3, 4)
    out = np.stack([((m[k, 0] * p[:, 0] + m[k, 1] * p[:, 1]) + m[k, 2] * p[:, 2]) + m[k, 3] for k in range(3)], axis=1)
    return out.reshape(-1, 9)


@pytest.mark.parametrize("name", ["cube", "elephant"])
def test_device_traversal_under_a_transformation(name, meshes):
    """the mesh under an affine map (does_A_contains_a_CC_of_B, Rigid_triangle_mesh_collision_detection.h:131-147): the device traversal with transformed
    datums and conservatively mapped node boxes (bvh.cuh::traverse_side<AffXf>) against the oracle on a tree BUILT over the rounded images"""
    from test_triangle_queries import rigid
    rng = np.random.default_rng(34)
    v, f = meshes[name]
    soup = v[f.astype(np.int64)].reshape(-1, 9)
    h = hc.HostTree(soup)
    for trial in range(3):
        m = rigid(rng, scale=rng.uniform(0.5, 2.0), shift=rng.uniform(-3, 3, 3)) if trial else np.array([[0, -1, 0, 0.5], [1, 0, 0, -0.25], [0, 0, 1, 2.0]], float)
        img = transformed_soup(m, soup)
        w = img.reshape(-1, 3)
        lo, hi = w.min(0), w.max(0)
        pts = np.concatenate([lo + rng.uniform(-0.1, 1.1, (600, 3)) * (hi - lo), w[rng.integers(0, len(w), 100)],
                              (w[rng.integers(0, len(w), 100)] + w[rng.integers(0, len(w), 100)]) / 2])
        want, _ = orc.OracleTree(soup=img).side_of_mesh(pts)
        got = h.side_of_mesh_xf(m, pts, DIRS)
        assert np.array_equal(got, want), (name, trial, np.flatnonzero(got != want)[:5])
        assert (want == 1).any() and (want == -1).any() and (want == 0).any()
