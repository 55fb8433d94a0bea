"""Pins the CPU oracle against the reference's own known-answer tests."""
import numpy as np
import pytest

import orc

SEG, RAY = 0, 1


def P(*a):
    return np.array(a, dtype=np.float64)


def test_correctness_triangle_kat(meshes):
    # AABB_tree/test/AABB_tree/aabb_correctness_triangle_test.cpp:51-108 (Data/data/meshes/tetrahedron.off)
    v, f = meshes["tetrahedron"]
    t = orc.OracleTree(v, f)
    pq = np.array([[-0.25, 0.251, 0.255, 0.25, 0.253, 0.256]])
    assert t.do_intersect(pq, SEG)[0] == 1
    assert t.count(pq, SEG)[0] == 1
    r = t.closest_point(np.array([[0.0, 0.0, 3.0]]))
    assert tuple(r["cp"][0]) == (0.0, 0.0, 1.0)
    # the same tree without the kd-tree hint gives the same point
    t2 = orc.OracleTree(v, f, accelerate=False)
    assert tuple(t2.closest_point(np.array([[0.0, 0.0, 3.0]]))["cp"][0]) == (0.0, 0.0, 1.0)


def test_triangle_3_kat():
    # AABB_tree/test/AABB_tree/aabb_test_triangle_3.cpp:27-46
    a, b, c, d = (1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0), (0, 0, 0)
    soup = np.array([a + b + c, a + b + d, a + d + c], dtype=np.float64)
    t = orc.OracleTree(soup=soup)
    assert t.count(np.array([a + b], dtype=np.float64), RAY)[0] == 3
    r = t.closest_point(np.array([[3.0, 2.0, 2.0]]))
    assert tuple(r["cp"][0]) == a
    assert r["sqd"][0] == 12.0


def test_face_graph_example_kat(meshes):
    # AABB_tree/examples/AABB_tree/AABB_face_graph_triangle_example.cpp:22-39 (unit tetrahedron p,q,r,s)
    v = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0, 0, 0]])
    f = np.array([[0, 1, 2], [1, 3, 2], [0, 3, 1], [0, 2, 3]], dtype=np.uint32)   # make_tetrahedron faces (any orientation)
    t = orc.OracleTree(v, f)
    assert t.count(np.array([[0.2, 0.2, 0.2, 1.0, 1.0, 1.0]]), SEG)[0] == 1


def test_empty_and_singleton_tree():
    # aabb_test_empty_tree.cpp:38-53, aabb_test_singleton_tree.cpp:44-67
    e = orc.OracleTree(soup=np.zeros((0, 9)))
    q = np.array([[0.1, 0.1, -1.0, 0.1, 0.1, 1.0]])
    assert e.size() == 0 and e.do_intersect(q, SEG)[0] == 0 and e.count(q, RAY)[0] == 0
    assert e.first_intersection(q)["prim"][0] == 0xFFFFFFFF
    with pytest.raises(RuntimeError):
        e.closest_point(np.zeros((1, 3)))
    s = orc.OracleTree(soup=np.array([[0, 0, 0, 1.0, 0, 0, 0, 1.0, 0]]))
    assert s.size() == 1
    assert s.do_intersect(q, SEG)[0] == 1 and s.count(q, RAY)[0] == 1
    fi = s.first_intersection(q)
    assert fi["prim"][0] == 0 and fi["kind"][0] == 1 and tuple(fi["pts"][0][:3]) == (0.1, 0.1, 0.0)
    assert tuple(s.closest_point(np.array([[0.25, 0.25, 5.0]]))["cp"][0]) == (0.25, 0.25, 0.0)
    np.testing.assert_array_equal(s.bbox(), [0, 0, 0, 1, 1, 0])


def test_ray_triangle_intersection_kats():
    # Intersections_3/test/Intersections_3/test_intersections_Ray_3.cpp:233-258
    def chk(ray, tri, kind, *pts):
        k, out = orc.tri_ray_intersection(np.array(tri, float), np.array(ray, float))
        assert k == kind, (ray, tri, k)
        assert orc.tri_do_intersect(np.array(tri, float), np.array(ray, float), RAY) == (1 if kind else 0)
        got = [tuple(out[i]) for i in range(kind)]
        if kind == 2:
            assert sorted(got) == sorted(pts)
        elif kind == 1:
            assert got[0] == pts[0]
    chk((0, 0, 0, 1, 1, 1), (-4, 2, 2, 3, 7, -1, 0, 9, -4), 0)
    chk((0, 0, 0, 1, 1, 1), (-1, -1, -1, 3, 7, -1, 0, 9, -4), 0)
    chk((0, 0, 0, 1, 1, 1), (0, 0, 0, 3, 7, -1, 0, 9, -4), 1, (0, 0, 0))
    chk((0, 0, 0, 1, 1, 1), (2, 2, 2, 3, 7, -1, 0, 9, -4), 1, (2, 2, 2))
    chk((0, 0, 0, 4, -2, 6), (2, -1, 3, 3, 7, -1, 0, 9, -4), 1, (2, -1, 3))
    chk((0, 0, 0, 2, -1, 3), (4, -2, 6, 3, 7, -2, 0, 8, -4), 1, (4, -2, 6))
    k, _ = orc.tri_ray_intersection(np.array((2, 4, 2, 5, -7, -1, 0, -9, 4), float), np.array((0, 0, 0, 1, 1, 1), float))
    assert k == 1
    chk((0, 0, 1, 1, 1, 1), (3, 3, 1, 7, 7, 1, 0, -9, -4), 2, (3, 3, 1), (7, 7, 1))
    chk((0, 0, 1, 1, 1, 1), (7, 7, 1, 3, 3, 1, 0, -9, -4), 2, (3, 3, 1), (7, 7, 1))
    chk((3, 3, 1, 7, 7, 1), (4, 4, 1, 2, 2, 1, 0, -9, -4), 2, (3, 3, 1), (4, 4, 1))


def test_segment_triangle_do_intersect_kats():
    # Intersections_3/test/Intersections_3/test_intersections_Segment_3.cpp:152-176 (do_intersect == has an intersection)
    cases = [
        ((1, 4, 7, 2, 3, 9), (-1, 9, 8, -5, 3, 4, -0, 8, 9), 0),
        ((4, 2, 7, 1, 3, 7), (1, 3, 7, 3, 9, -1, -7, 6, 3), 1),
        ((1, 3, 9, 5, 1, 7), (3, 2, 8, -1, 9, 7, -2, 2, 8), 1),
        ((3, 2, 7, 1, -2, -9), (3, 2, 7, 1, -2, -9, 4, 3, 9), 1),
        ((3, 2, 7, 1, -2, -9), (5, 6, 23, 1, -2, -9, 4, 3, 9), 1),
        ((3, 2, 7, 1, -2, -9), (5, 6, 23, -1, -6, -25, 4, 3, 9), 1),
        ((1, 3, 9, 5, 1, 7), (4, 0, 6, 2, 4, 10, -2, 2, 8), 1),
        ((1, 3, 7, 5, 1, 7), (4, 0, 7, 2, 4, 7, -2, 2, 7), 1),
        ((1, 3, 7, 5, 1, 7), (4, 0, 7, 2, 4, 7, 1, 2, 7), 1),
    ]
    for seg, tri, want in cases:
        assert orc.tri_do_intersect(np.array(tri, float), np.array(seg, float), SEG) == want, (seg, tri)


@pytest.mark.parametrize("name", ["cube", "tetrahedron", "coverrear", "pinion_small"])
def test_tree_vs_naive_intersections(meshes, name):
    # AABB_tree/test/AABB_tree/aabb_naive_vs_tree_triangle_test.cpp:33-50 + AABB_test_util.h:572-676
    v, f = meshes[name]
    t = orc.OracleTree(v, f)
    rng = np.random.default_rng(23)
    bb = t.bbox()
    n = 300
    q = rng.uniform(bb[:3], bb[3:], size=(n, 2, 3)).reshape(n, 6)   # two random points in the bbox (AABB_test_util.h:41-48)
    for kind in (SEG, RAY):
        c_tree, c_naive = t.count(q, kind), t.count(q, kind, naive=True)
        np.testing.assert_array_equal(c_tree, c_naive)
        np.testing.assert_array_equal(t.do_intersect(q, kind), (c_naive > 0).astype(np.uint8))


@pytest.mark.parametrize("name", ["cube", "coverrear", "pinion_small"])
def test_tree_vs_naive_distance(meshes, name):
    # aabb_naive_vs_tree_distance_triangle_test.cpp:33-57 + AABB_test_util.h:755-830 — the reference only compares
    # squared distances (tolerance 1e-7); the restatement is held to: same point up to exact ties.
    v, f = meshes[name]
    t = orc.OracleTree(v, f)
    rng = np.random.default_rng(23)
    bb = t.bbox()
    q = rng.uniform(bb[:3] - 0.3 * (bb[3:] - bb[:3]), bb[3:] + 0.3 * (bb[3:] - bb[:3]), size=(300, 3))
    a, b = t.closest_point(q), t.naive_closest(q)
    d_tree = ((a["cp"] - q) ** 2).sum(1)
    d_naive = ((b["cp"] - q) ** 2).sum(1)
    assert np.abs(d_tree - d_naive).max() <= 1e-7
    same = (a["prim"] == b["prim"])
    assert np.all(same | (b["ntie"] > 1))
    for i in np.nonzero(~same)[0]:   # mismatching id must be an exact-distance tie
        assert orc.compare_distance(q[i], a["cp"][i], b["cp"][i]) == 0
    assert np.all((a["cp"] == b["cp"]).all(1) | (b["ntie"] > 1))


def test_first_intersection_is_argmin_of_all(meshes):
    # aabb_test_ray_intersection.cpp:38-65,87-140: rays between points of a sphere about the bbox centre
    v, f = meshes["coverrear"]
    t = orc.OracleTree(v, f)
    rng = np.random.default_rng(23)
    bb = t.bbox()
    c, r = 0.5 * (bb[:3] + bb[3:]), 0.5 * (bb[3:] - bb[:3]).max()
    u = rng.normal(size=(400, 2, 3))
    u /= np.linalg.norm(u, axis=2, keepdims=True)
    rays = (c + r * u).reshape(-1, 6)
    a, b = t.first_intersection(rays), t.naive_first_intersection(rays)
    hit = b["prim"] != 0xFFFFFFFF
    assert hit.sum() > 50
    np.testing.assert_array_equal(a["prim"] != 0xFFFFFFFF, hit)
    np.testing.assert_array_equal(a["t"][hit], b["t"][hit])
    same = a["prim"] == b["prim"]
    assert np.all(same | (b["ntie"] > 1))


def test_segment_triangle_intersection_kats():
    # Intersections_3/test/Intersections_3/test_intersections_Segment_3.cpp:152-176: the intersection OBJECTS (Segment_3_Triangle_3_intersection.h)
    def chk(seg, tri, kind, *pts):
        k, out = orc.tri_seg_intersection(np.array(tri, float), np.array(seg, float))
        assert k == kind, (seg, tri, k)
        got = [tuple(out[i]) for i in range(k)]
        assert sorted(got) == sorted(pts), (seg, tri, got)
    chk((1, 4, 7, 2, 3, 9), (-1, 9, 8, -5, 3, 4, -0, 8, 9), 0)
    chk((4, 2, 7, 1, 3, 7), (1, 3, 7, 3, 9, -1, -7, 6, 3), 1, (1, 3, 7))
    chk((1, 3, 9, 5, 1, 7), (3, 2, 8, -1, 9, 7, -2, 2, 8), 1, (3, 2, 8))
    chk((3, 2, 7, 1, -2, -9), (3, 2, 7, 1, -2, -9, 4, 3, 9), 2, (3, 2, 7), (1, -2, -9))
    chk((3, 2, 7, 1, -2, -9), (5, 6, 23, 1, -2, -9, 4, 3, 9), 2, (3, 2, 7), (1, -2, -9))
    chk((3, 2, 7, 1, -2, -9), (5, 6, 23, -1, -6, -25, 4, 3, 9), 2, (3, 2, 7), (1, -2, -9))
    chk((1, 3, 9, 5, 1, 7), (4, 0, 6, 2, 4, 10, -2, 2, 8), 2, (1, 3, 9), (3, 2, 8))
    for tri in ((4, 0, 7, 2, 4, 7, -2, 2, 7), (4, 0, 7, 2, 4, 7, 1, 2, 7)):                     # :172-173 only ask for a Segment_3
        assert orc.tri_seg_intersection(np.array(tri, float), np.array((1, 3, 7, 5, 1, 7), float))[0] == 2
