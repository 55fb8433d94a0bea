"""Trees over Segment_3 primitives (AABB_segment_primitive_3, AABB_halfedge_graph_segment_primitive): closest_point /
squared_distance and the id queries with Plane_3 and Triangle_3 — what AABB_tree/examples/AABB_tree/AABB_segment_3_example.cpp and
AABB_halfedge_graph_edge_example.cpp ask of such a tree.

1. the oracle (oracle/tree.hpp with Mesh::segments): the reference's example as a known answer, tree == brute force (the reference's own differential test,
   AABB_tree/test/AABB_tree/aabb_naive_vs_tree_distance_segment_test.cpp), and python rationals for the three leaf routines;
2. the device code (bvh.cuh::traverse_closest<true>, traverse_triangle<.., true>, traverse_plane on the (source, target, target) records, compiled for the
   host) against the oracle.
"""
from fractions import Fraction as Fr

import numpy as np
import pytest

import hc
import orc
from test_oracle_predicates import line_tri_fr

PLANE, TRIANGLE = 3, 4


def records(segs):
    """the (source, target, target) records the device tree stores for Segment_3 primitives"""
    s = np.asarray(segs, float).reshape(-1, 6)
    return np.concatenate([s, s[:, 3:]], axis=1)


def mesh_edges(v, f):
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
    e = np.unique(np.sort(e, axis=1), axis=0)
    return np.concatenate([v[e[:, 0]], v[e[:, 1]]], axis=1)


def test_reference_example():
    """AABB_segment_3_example.cpp:26-58: three segments; the plane (a, b, d) and the triangle (a, b, c) meet all three; the point (2, 2, 2) is equally far from
    (a, b) and (a, c)"""
    a, b, c, d = [1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0, 0, 0]
    segs = np.array([a + b, a + c, c + d])
    o = orc.OracleTree(segments=segs)
    assert o.size() == 3
    plane = np.array([[0.0, 0.0, 1.0, 0.0]])                                # through a, b, d: z = 0 (Plane_3(a, b, d) up to orientation and scale)
    assert o.count(plane, PLANE)[0] == 3
    assert o.count(np.array([a + b + c]), TRIANGLE)[0] == 3
    r = o.closest_point(np.array([[2.0, 2.0, 2.0]]))
    assert tuple(r["cp"][0]) in ((0.5, 0.5, 0.0), (0.5, 0.0, 0.5)) and r["sqd"][0] == 8.5
    h = hc.HostTree(records(segs))
    g = h.closest_point_segments(np.array([[2.0, 2.0, 2.0]]))
    assert g["sqd"][0] == 8.5 and g["tie"][0] == 1 and g["prim"][0] == 0   # the tie is reported, the smallest id wins
    assert h.count_line_plane(plane, PLANE)[0] == 3 and h.count_triangle_segments(np.array([a + b + c]))[0] == 3


def closest_on_segment_fr(a, b, q):
    a, b, q = [Fr(float(x)) for x in a], [Fr(float(x)) for x in b], [Fr(float(x)) for x in q]
    d = [y - x for x, y in zip(a, b)]
    t = sum((qi - ai) * di for qi, ai, di in zip(q, a, d)) / sum(di * di for di in d)
    t = min(max(t, 0), 1)
    p = [ai + t * di for ai, di in zip(a, d)]
    return sum((pi - qi) ** 2 for pi, qi in zip(p, q))


def test_oracle_segment_tree(meshes):
    rng = np.random.default_rng(51)
    v, f = meshes["elephant"]
    segs = mesh_edges(v, f)[::7]                                            # every 7th edge of the mesh: ~1,200 segments
    o = orc.OracleTree(segments=segs)
    lo, hi = v.min(0), v.max(0)
    pts = np.concatenate([lo + rng.uniform(-0.2, 1.2, (1500, 3)) * (hi - lo), segs[rng.integers(0, len(segs), 100), :3],
                          (segs[rng.integers(0, len(segs), 100), :3] + segs[rng.integers(0, len(segs), 100), 3:]) / 2])
    r, nv = o.closest_point(pts), o.naive_closest(pts)
    same = r["prim"] == nv["prim"]
    assert np.array_equal(r["cp"][same], nv["cp"][same]) and (same | (nv["ntie"] > 1)).all()        # tree == brute force up to exact ties
    for i in range(0, len(pts), 40):                                        # the squared distance is the exact minimum up to the rounding of the projection
        want = min(closest_on_segment_fr(s[:3], s[3:], pts[i]) for s in segs)
        assert abs(Fr(float(r["sqd"][i])) - want) <= want * Fr(1, 10**12) + Fr(1, 10**300)
    nrm = rng.normal(size=(400, 3)); x = lo + rng.uniform(0, 1, (400, 3)) * (hi - lo)
    planes = np.concatenate([nrm, -np.sum(nrm * x, axis=1, keepdims=True)], axis=1)
    planes[::4, :3] = np.eye(3)[rng.integers(0, 3, 100)]; planes[::4, 3] = -segs[rng.integers(0, len(segs), 100), 0]      # axis planes through segment sources' x
    ctr = lo + rng.uniform(0, 1, (400, 3)) * (hi - lo)
    tris = np.concatenate([ctr, ctr + rng.uniform(-0.2, 0.2, (400, 3)) * (hi - lo), ctr + rng.uniform(-0.2, 0.2, (400, 3)) * (hi - lo)], axis=1)
    tris[::5] = v[f[rng.integers(0, len(f), 80)]].reshape(-1, 9)            # facets of the mesh: their edges lie in them
    for kind, q in ((PLANE, planes), (TRIANGLE, tris)):
        want = o.count(q, kind)
        assert np.array_equal(want, o.count(q, kind, naive=True))
        assert want.max() > 2 and (want == 0).any()
    for i in range(0, len(tris), 8):                                        # Triangle_3 query x Segment_3 primitive against the rational segment / triangle test
        want = sum(line_tri_fr(tris[i], s[:3], s[3:], 0) for s in segs[::3])
        assert o.__class__(segments=segs[::3]).count(tris[i:i + 1], TRIANGLE)[0] == want


@pytest.mark.parametrize("name", ["cube", "elephant"])
def test_device_code_against_the_oracle(name, meshes):
    rng = np.random.default_rng(52)
    v, f = meshes[name]
    segs = mesh_edges(v, f)
    if name == "elephant":
        segs = segs[::5]
    o, h = orc.OracleTree(segments=segs), hc.HostTree(records(segs))
    lo, hi = v.min(0), v.max(0)
    pts = np.concatenate([lo + rng.uniform(-0.2, 1.2, (1200, 3)) * (hi - lo), segs[rng.integers(0, len(segs), 100), :3],
                          (segs[rng.integers(0, len(segs), 100), :3] + segs[rng.integers(0, len(segs), 100), 3:]) / 2, (lo + hi).reshape(1, 3) / 2])
    want, got = o.closest_point(pts), h.closest_point_segments(pts)
    assert np.array_equal(got["sqd"], want["sqd"])
    same = got["prim"] == want["prim"]
    assert np.array_equal(got["cp"][same], want["cp"][same])
    assert (same | (got["tie"] == 1)).all()                                 # a different primitive only where the device reports an exact tie
    nrm = rng.normal(size=(300, 3)); x = lo + rng.uniform(0, 1, (300, 3)) * (hi - lo)
    planes = np.concatenate([nrm, -np.sum(nrm * x, axis=1, keepdims=True)], axis=1)
    planes[::3, :3] = np.eye(3)[rng.integers(0, 3, 100)]; planes[::3, 3] = -segs[rng.integers(0, len(segs), 100), 0]
    ctr = lo + rng.uniform(0, 1, (300, 3)) * (hi - lo)
    tris = np.concatenate([ctr, ctr + rng.uniform(-0.3, 0.3, (300, 3)) * (hi - lo), ctr + rng.uniform(-0.3, 0.3, (300, 3)) * (hi - lo)], axis=1)
    tris[::5] = v[f[rng.integers(0, len(f), 60)]].reshape(-1, 9)
    assert np.array_equal(h.count_line_plane(planes, PLANE), o.count(planes, PLANE))
    assert np.array_equal(h.count_triangle_segments(tris), o.count(tris, TRIANGLE))
