"""Triangle_3 queries and tree-vs-tree do_intersect (`tree.do_intersect(Triangle_3)`, `tree.do_intersect(other_tree)`,
AABB_traits_3.h:673-690; Do_intersect_3(Triangle_3, Triangle_3) = Intersections_3/internal/Triangle_3_Triangle_3_do_intersect.h).

1. pins the oracle's restatement (oracle/kernel.hpp::do_intersect_tri_tri) with the reference's own known answers
   (tests/golden/tri_tri_kats.json, extracted from Intersections_3/test/Intersections_3/triangle_3_triangle_3_intersection.cpp by
   tests/golden/make_tri_tri_kats.py) and with an independent exact formulation: two closed triangles meet iff an edge of one meets the other,
   decided by the segment / triangle test in python rationals (tests/test_oracle_predicates.py::line_tri_fr);
2. runs the DEVICE code (geom.cuh::tri_tri_do_intersect, bvh.cuh::traverse_triangle compiled for the host) against the oracle, predicate by predicate
   and through whole traversals, on random, touching, coplanar and scaled inputs.
"""
import json
import os

import numpy as np
import pytest

import hc
import orc
from test_oracle_predicates import line_tri_fr

TRIANGLE = 4
KATS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tri_tri_kats.json")))["kats"]


def flat(t):
    a, b, c = t[0:3], t[3:6], t[6:9]
    return not np.any(np.cross(b - a, c - a))


def triangles_meet_fr(t1, t2):
    """exact: some edge of t1 meets t2 or some edge of t2 meets t1 (closed sets; also covers containment in a common plane)"""
    for s, t in ((t1, t2), (t2, t1)):
        for i, j in ((0, 1), (1, 2), (2, 0)):
            if line_tri_fr(t, s[3 * i:3 * i + 3], s[3 * j:3 * j + 3], 0):
                return True
    return False


def test_reference_known_answers():
    assert len(KATS) >= 40 and sum(not k["intersect"] for k in KATS) >= 4
    for k in KATS:
        t1, t2 = np.array(k["t1"]), np.array(k["t2"])
        for a, b in ((t1, t2), (t2, t1)):                                  # the reference test asserts both argument orders
            assert bool(orc.tri_tri_do_intersect(a, b)) == k["intersect"], ("oracle", k["line"])
            assert bool(hc.tri_tri_do_intersect(a, b)) == k["intersect"], ("device code", k["line"])
        assert triangles_meet_fr(t1, t2) == k["intersect"], ("rationals", k["line"])


def pair_cases(rng, n):
    """pairs of non-flat triangles: general position, small lattices (touching / coplanar / shared vertices and edges are frequent), and pairs built
    to share a plane, an edge or a vertex"""
    out = []
    while len(out) < n:
        m = len(out) % 6
        if m == 0:
            t1, t2 = rng.uniform(-1, 1, 9), rng.uniform(-1, 1, 9)
        elif m in (1, 2):
            t1, t2 = rng.integers(-2, 3, 9).astype(float), rng.integers(-2, 3, 9).astype(float)
        elif m == 3:                                                        # coplanar lattice pairs (z = const, or a common vertical plane)
            t1, t2 = rng.integers(-3, 4, 9).astype(float), rng.integers(-3, 4, 9).astype(float)
            ax = int(rng.integers(0, 3)); c = float(rng.integers(-1, 2))
            t1[ax::3] = c; t2[ax::3] = c
        elif m == 4:                                                        # shared edge
            t1 = rng.integers(-3, 4, 9).astype(float); t2 = rng.integers(-3, 4, 9).astype(float)
            t2[0:6] = t1[3:9]
        else:                                                               # a vertex of one on an edge of the other
            t1 = rng.integers(-4, 5, 9).astype(float) * 2; t2 = rng.integers(-4, 5, 9).astype(float) * 2
            t2[0:3] = (t1[0:3] + t1[3:6]) / 2
        if flat(t1) or flat(t2):
            continue
        out.append((t1, t2))
    return out


def test_oracle_against_rationals():
    rng = np.random.default_rng(41)
    n_true = 0
    for t1, t2 in pair_cases(rng, 3000):
        want = triangles_meet_fr(t1, t2)
        n_true += want
        assert bool(orc.tri_tri_do_intersect(t1, t2)) == want, (t1, t2)
        assert bool(orc.tri_tri_do_intersect(t2, t1)) == want, (t2, t1)
    assert 600 < n_true < 2700


def test_device_predicate_against_the_oracle():
    rng = np.random.default_rng(42)
    n_true = 0
    for scale in (1.0, 1e-3, 37.0, 1e-160, 1e150):                          # (outside [1e-140, 1e140] the static filters must hand over to the exact stage)
        for t1, t2 in pair_cases(rng, 1500):
            a, b = t1 * scale, t2 * scale
            want = orc.tri_tri_do_intersect(a, b)
            n_true += want
            assert hc.tri_tri_do_intersect(a, b) == want, (scale, t1, t2)
            assert hc.tri_tri_do_intersect(b, a) == want, (scale, t2, t1)
    assert n_true > 1000
    # nearly touching pairs: one vertex of t2 moved by a few ulps across the plane / an edge line of t1
    for _ in range(2000):
        t1 = rng.uniform(-1, 1, 9)
        w = rng.dirichlet(np.ones(3)) if rng.integers(0, 2) else np.append(rng.dirichlet(np.ones(2)), 0.0)
        x = w[0] * t1[0:3] + w[1] * t1[3:6] + w[2] * t1[6:9]
        x = np.nextafter(x, x + rng.choice([-1.0, 1.0], 3) * rng.integers(0, 3, 3))
        t2 = np.concatenate([x, x + rng.uniform(-1, 1, 3), x + rng.uniform(-1, 1, 3)])
        assert hc.tri_tri_do_intersect(t1, t2) == orc.tri_tri_do_intersect(t1, t2), (t1, t2)
        assert hc.tri_tri_do_intersect(t2, t1) == orc.tri_tri_do_intersect(t2, t1), (t2, t1)


def mesh_triangles(rng, v, f, n):
    """query triangles around a mesh: random, copies of facets (coplanar with their primitive), facets moved along their plane, slivers through vertices"""
    lo, hi = v.min(0), v.max(0)
    ctr, ext = (lo + hi) / 2, (hi - lo).max()
    out = []
    while len(out) < n:
        m = len(out) % 5
        if m == 0:
            p = ctr + rng.uniform(-0.6, 0.6, 3) * ext
            t = np.concatenate([p, p + rng.uniform(-0.2, 0.2, 3) * ext, p + rng.uniform(-0.2, 0.2, 3) * ext])
        elif m == 1:
            t = v[f[rng.integers(0, len(f))]].ravel().copy()
        elif m == 2:                                                        # a facet shifted inside its own plane
            a, b, c = v[f[rng.integers(0, len(f))]]
            d = 0.5 * (b - a) + 0.25 * (c - a)
            t = np.concatenate([a + d, b + d, c + d])
        elif m == 3:                                                        # through two mesh vertices and a random point
            a, b = v[rng.integers(0, len(v), 2)]
            t = np.concatenate([a, b, ctr + rng.uniform(-1, 1, 3) * ext])
        else:                                                               # far outside
            p = ctr + 3 * ext * np.array([1.0, 1.0, 1.0])
            t = np.concatenate([p, p + rng.uniform(0, 1, 3), p + rng.uniform(0, 1, 3)])
        if not flat(t):
            out.append(t)
    return np.array(out)


@pytest.mark.parametrize("mesh", ["cube", "pinion_small", "lattice"])
def test_device_traversal_against_the_oracle(mesh, meshes):
    rng = np.random.default_rng(43)
    if mesh == "lattice":                                                   # two crossing 5 x 5 sheets of unit right triangles
        g = np.arange(5, dtype=np.float64)
        v = np.array([[x, y, 0.0] for y in g for x in g] + [[x, 2.0, z] for z in g for x in g])
        quad = np.array([[j * 5 + i, j * 5 + i + 1, j * 5 + i + 6, j * 5 + i + 5] for j in range(4) for i in range(4)])
        one = np.concatenate([quad[:, [0, 1, 2]], quad[:, [0, 2, 3]]])
        f = np.concatenate([one, one + 25]).astype(np.uint32)
    else:
        v, f = meshes[mesh]
    soup = v[f].reshape(-1, 9)
    q = mesh_triangles(rng, v, f, 300 if mesh != "pinion_small" else 150)
    o, h = orc.OracleTree(v, f), hc.HostTree(soup)
    want = o.count(q, TRIANGLE)
    assert np.array_equal(want, o.count(q, TRIANGLE, naive=True))           # traversal == brute force, inside the oracle
    assert np.array_equal(o.do_intersect(q, TRIANGLE), (want > 0).astype(np.uint8))
    assert np.array_equal(h.intersect_triangle(q, True), want)
    assert np.array_equal(h.intersect_triangle(q, False), (want > 0).astype(np.uint32))
    assert want.max() > 1 and (want == 0).any()
    off, ids = o.all_intersected_primitives(q, TRIANGLE)
    anyp = o.any_intersected_primitive(q, TRIANGLE)
    for i in range(len(q)):
        assert (anyp[i] == 0xFFFFFFFF) == (want[i] == 0)
        if want[i]:
            assert anyp[i] in ids[int(off[i]):int(off[i + 1])]


def test_oracle_tree_against_tree(meshes):
    v, f = meshes["cube"]
    a = orc.OracleTree(v, f)
    ext = (v.max(0) - v.min(0)).max()
    for shift, want in ((0.0, True), (0.5 * ext, True), (1.0 * ext, True), (1.0001 * ext + 1e-9, False), (3.0 * ext, False)):   # touching faces still intersect
        b = orc.OracleTree(v + np.array([shift, 0.0, 0.0]), f)
        assert a.do_intersect_tree(b) == want, shift
        assert b.do_intersect_tree(a) == want, shift


def rigid(rng, scale=1.0, shift=(0.0, 0.0, 0.0)):
    """3 x 4 matrix of a rotation about a random axis, a uniform scaling and a translation"""
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    ang = rng.uniform(0, 2 * np.pi)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    return np.hstack([scale * R, np.asarray(shift, float).reshape(3, 1)])


def test_transformed_primitives_and_boxes():
    """bvh.cuh::AffXf: the image of a primitive is the reference's Aff_transformation_3::transform (oracle, evaluated left to right in binary64), and the
    conservative image of a node box contains the rounded images of everything inside the box"""
    rng = np.random.default_rng(45)
    for it in range(300):
        m = rigid(rng, scale=rng.uniform(0.5, 2.0), shift=rng.uniform(-5, 5, 3)) if it % 3 else np.hstack([rng.uniform(-2, 2, (3, 3)), rng.uniform(-5, 5, (3, 1))])
        lo = rng.uniform(-3, 3, 3).astype(np.float32)
        hi = (lo + rng.uniform(0, 2, 3).astype(np.float32)).astype(np.float32)
        for _ in range(20):
            w = rng.uniform(0, 1, (3, 3))
            if rng.integers(0, 2):
                w = np.round(w)                                             # vertices at the corners of the box
            tri = (lo + w * (hi.astype(np.float64) - lo)).ravel()
            img = hc.xf_apply(m, tri)
            p = tri.reshape(3, 3)
            want = np.array([[m[k, 0] * p[i, 0] + m[k, 1] * p[i, 1] + m[k, 2] * p[i, 2] + m[k, 3] for k in range(3)] for i in range(3)]).ravel()
            assert np.array_equal(img, want)
            assert hc.xf_box_meets(m, lo, hi, img)                          # the image of a triangle inside the box meets the image of the box
        far = hc.xf_apply(m, (hi + np.array([50.0, 50.0, 50.0])).repeat(3).reshape(3, 3).T.ravel() + rng.uniform(0, 1, 9))
        if it % 3:                                                          # (a rigid map keeps distances: something 50 away stays away)
            assert not hc.xf_box_meets(m, lo, hi, far)


def test_oracle_transformed_tree_against_tree(meshes):
    v, f = meshes["cube"]
    a, b = orc.OracleTree(v, f), orc.OracleTree(v, f)
    rng = np.random.default_rng(46)
    ext = (v.max(0) - v.min(0)).max()
    n_true = 0
    for it in range(60):
        xa = rigid(rng, shift=rng.uniform(-1, 1, 3) * ext) if it % 2 else None
        xb = rigid(rng, scale=rng.uniform(0.3, 1.5), shift=rng.uniform(-1.5, 1.5, 3) * ext)
        want = a.do_intersect_tree_transformed(xa, b, xb, naive=True)
        n_true += want
        assert a.do_intersect_tree_transformed(xa, b, xb) == want, it
        assert b.do_intersect_tree_transformed(xb, a, xa) == want, it
    assert 10 < n_true < 55
    ident = np.hstack([np.eye(3), np.zeros((3, 1))])
    assert a.do_intersect_tree_transformed(ident, b, None) == a.do_intersect_tree(b)


def test_device_tree_against_tree_under_transformations(meshes):
    """the traversal the device kernel runs (traverse_triangle<kAnyId, AffXf> per primitive of the other tree, compiled for the host) against the oracle's
    brute force over all pairs of transformed primitives; the witness pair is re-checked on the transformed triangles"""
    v, f = meshes["cube"]
    pv, pf = meshes["pinion_small"]
    rng = np.random.default_rng(49)
    ident = np.hstack([np.eye(3), np.zeros((3, 1))])
    for (va, fa), (vb, fb), trials in (((v, f), (v, f), 30), ((pv, pf), (v * (0.3 * (pv.max() - pv.min()) / (v.max() - v.min())) + pv.mean(0), f), 12)):
        sa, sb = va[fa].reshape(-1, 9), vb[fb].reshape(-1, 9)
        ha, hb = hc.HostTree(sa), hc.HostTree(sb)
        oa, ob = orc.OracleTree(va, fa), orc.OracleTree(vb, fb)
        ext = (va.max(0) - va.min(0)).max()
        n_true = 0
        for it in range(trials):
            xa = rigid(rng, shift=rng.uniform(-0.5, 0.5, 3) * ext) if it % 2 else ident
            xb = rigid(rng, scale=rng.uniform(0.5, 1.5), shift=rng.uniform(-1.0, 1.0, 3) * ext)
            want = oa.do_intersect_tree_transformed(xa, ob, xb, naive=True)
            n_true += want
            hit, ia, ib = ha.do_intersect_tree_xf(xa, hb, xb)
            assert hit == want, it
            if hit:
                assert orc.tri_tri_do_intersect(hc.xf_apply(xb, sb[ib]), hc.xf_apply(xa, sa[ia])) == 1
        assert 0 < n_true < trials
