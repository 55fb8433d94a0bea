"""Checks the per-thread logic of the CUDA kernels (device headers compiled for the host, tests/hostcheck) against the
oracle on CPU: predicates on random and degenerate inputs, bit-exact constructions, LBVH topology, and whole queries."""
import numpy as np
import pytest

import hc
import orc

SEG, RAY = 0, 1


def soup_of(v, f):
    return v[f.astype(np.int64)].reshape(-1, 9)


def lattice(rng, shape, lo=-2, hi=3):
    return rng.integers(lo, hi, size=shape).astype(float)


def test_struct_sizes():
    assert hc.lib().hc_sizeof_node() == 64 and hc.lib().hc_sizeof_tri() == 96


def test_predicates_match_oracle():
    rng = np.random.default_rng(11)
    for it in range(4000):
        mode = it % 4
        pts = rng.uniform(-10, 10, (5, 3)) if mode == 0 else lattice(rng, (5, 3))
        if mode == 2:
            w = rng.uniform(-1, 2, 3); w /= w.sum()
            pts = rng.uniform(-10, 10, (5, 3)); pts[3] = w @ pts[:3]
        if mode == 3:
            pts = lattice(rng, (5, 3)) + 4096.0
            pts[3] = np.nextafter(pts[3], pts[3] + rng.choice([-1, 1], 3))
        p4 = pts[:4].ravel()
        assert hc.call("hc_orient3", p4) == orc.orientation(*pts[:4])
        assert hc.call("hc_orient3_exact", p4) == orc.orientation(*pts[:4], exact=True)
        assert hc.call("hc_orient3_dir", pts.ravel()) == orc.orientation_dir(*pts)
        assert hc.call("hc_coplanar_orient", pts[:3].ravel()) == orc.coplanar_orientation(*pts[:3])
        p, q = pts[0], pts[1]
        r = [pts[2], 2 * p - q, np.nextafter(q, q + 1)][it % 3]
        tri = np.concatenate([p, q, r])
        assert hc.call("hc_cmp_dist", tri) == orc.compare_distance(p, q, r)
        assert hc.call("hc_cmp_dist_exact", tri) == orc.compare_distance(p, q, r, exact=True)


def test_compare_distance_double_double_stage_is_exact():
    """the double-double stage of Compare_distance_3 (predicates.cuh: cmp_dist_dd) decides the cases the static filter cannot:
    candidates a few ulps apart (shared-edge projections), mirrored points (exact ties -> expansion stage), huge offsets, tiny scales"""
    rng = np.random.default_rng(77)
    nz = 0
    for it in range(6000):
        scale = [1.0, 4096.0, 1e-6, 1e8][it % 4]
        p = rng.uniform(-1, 1, 3) * scale + (it % 3) * scale
        q = rng.uniform(-1, 1, 3) * scale + (it % 3) * scale
        mode = (it // 4) % 5
        if mode == 0:      # r = q moved by -2..2 ulps per coordinate
            r = q.copy()
            for _ in range(2):
                r = np.nextafter(r, r + rng.choice([-1.0, 0.0, 1.0], 3) * scale)
        elif mode == 1:    # mirror image of q about p, then a few ulps off
            r = 2 * p - q
            if it % 2:
                r = np.nextafter(r, r + rng.choice([-1.0, 0.0, 1.0], 3) * scale)
        elif mode == 2:    # same distance by a permutation / sign flip of the offset: exact ties when the arithmetic is exact
            d = np.round((q - p) / scale * 1024) / 1024 * scale
            q = p + d
            r = p + rng.permutation(d) * rng.choice([-1.0, 1.0], 3)
        elif mode == 3:    # nearly equal distances in unrelated directions
            d = q - p
            e = rng.normal(size=3); e *= np.linalg.norm(d) / np.linalg.norm(e)
            r = p + e
        else:
            r = rng.uniform(-1, 1, 3) * scale
        want = orc.compare_distance(p, q, r, exact=True)
        assert hc.call("hc_cmp_dist", np.concatenate([p, q, r])) == want
        assert orc.compare_distance(p, q, r) == want
        nz += want != 0
    assert nz > 3000


def test_triangle_tests_and_constructions_match_oracle_bitwise():
    rng = np.random.default_rng(12)
    n_hit = n_seg = 0
    for it in range(10000):
        if it % 4 == 3:   # everything in one axis-parallel or oblique lattice plane: the coplanar case analyses
            tri, pq = lattice(rng, 9, -3, 4), lattice(rng, 6, -3, 4)
            ax = it % 3
            tri[ax::3] = 1.0; pq[ax::3] = 1.0
            if it % 8 == 7:   # oblique plane x + y + z = 3
                tri[2::3] = 3 - tri[0::3] - tri[1::3]; pq[2::3] = 3 - pq[0::3] - pq[1::3]
        elif it % 3 == 0:
            tri, pq = rng.uniform(-1, 1, 9), rng.uniform(-1, 1, 6)
        elif it % 3 == 1:
            tri, pq = lattice(rng, 9), lattice(rng, 6)
        else:   # ray aimed at a point of the triangle: exercises the hit construction
            tri = rng.uniform(-1, 1, 9)
            w = rng.dirichlet([1, 1, 1])
            target = w @ tri.reshape(3, 3)
            s = rng.uniform(-3, 3, 3)
            pq = np.concatenate([s, s + (target - s) * rng.uniform(0.2, 3)])
        a, b, c = tri[:3], tri[3:6], tri[6:]
        if not np.any(np.cross(b - a, c - a)) or np.all(pq[:3] == pq[3:]):
            continue
        for is_ray in (0, 1):
            assert hc.tri_do_intersect(tri, pq, is_ray) == orc.tri_do_intersect(tri, pq, is_ray), (tri, pq, is_ray)
        k, pts, t = hc.tri_ray_intersection(tri, pq)
        ko, ptso = orc.tri_ray_intersection(tri, pq)
        assert k == ko, (tri, pq)
        if k:
            n_hit += 1
            n_seg += k == 2
            np.testing.assert_array_equal(pts[:k], ptso[:k])     # bit-exact
        q = rng.uniform(-2, 2, 3) if it % 3 != 1 else lattice(rng, 3)
        np.testing.assert_array_equal(hc.closest_on_triangle(tri, q), orc.project_on_triangle(tri, q))
    assert n_hit > 1500 and n_seg > 300, (n_hit, n_seg)
    # degenerate triangle (documented UB for the tree, still restated): longest-edge segment projection
    for tri in ([0, 0, 0, 1.0, 1, 1, 2, 2, 2], [0, 0, 0, 0, 0, 0, 1.0, 0, 0], [1.0, 2, 3, 1, 2, 3, 1, 2, 3]):
        for q in ([0.3, 0.1, 5.0], [-4, 0, 0], [9.0, 9, 9]):
            np.testing.assert_array_equal(hc.closest_on_triangle(tri, q), orc.project_on_triangle(tri, q))


def test_node_test_is_conservative():
    rng = np.random.default_rng(13)
    n_true = 0
    for it in range(6000):
        if it % 2:
            lo = lattice(rng, 3, -2, 1); hi = lo + rng.integers(0, 3, 3)
            pq = lattice(rng, 6, -3, 4)
        else:
            lo = rng.uniform(-1, 0, 3); hi = lo + rng.uniform(0.01, 1, 3)
            corner = np.where(rng.integers(0, 2, 3) == 1, hi, lo)
            p = rng.uniform(-3, 3, 3)
            pq = np.concatenate([p, p + (corner - p) * rng.choice([0.5, 1.0, 2.0])])
        if np.all(pq[:3] == pq[3:]):
            continue
        box = np.concatenate([lo, hi])
        for is_ray in (0, 1):
            exact = orc.bbox_do_intersect(box, pq, is_ray)
            n_true += exact
            if exact:
                assert hc.line_box(box, pq, is_ray) == 1, (box, pq, is_ray)
    assert n_true > 1000
    # terrain-like magnitudes: unit boxes near coordinate 4096 (binary32 spacing 4.9e-4), sources far away, lines aimed
    # exactly at box corners / edges (grazing contact must never be rejected by the binary32 node test)
    n_true = 0
    for it in range(6000):
        lo = np.concatenate([rng.integers(0, 4096, 2).astype(float), rng.uniform(-40, 40, 1)])
        hi = lo + np.array([1.0, 1.0, rng.uniform(0.0, 0.5)])
        corner = np.where(rng.integers(0, 2, 3) == 1, hi, lo)
        p = np.array([0.0, 0.0, 600.0]) if it % 2 else rng.uniform(-5000, 5000, 3)
        target = corner if it % 3 else lo + rng.uniform(0, 1, 3) * (hi - lo)
        pq = np.concatenate([p, p + (target - p) * rng.choice([0.25, 1.0, 3.0])])
        box = np.concatenate([lo, hi])
        for is_ray in (0, 1):
            if orc.bbox_do_intersect(box, pq, is_ray):
                n_true += 1
                assert hc.line_box(box, pq, is_ray) == 1, (box, pq, is_ray)
    assert n_true > 3000


def test_point_box_distance_is_a_lower_bound():
    from fractions import Fraction as Fr
    rng = np.random.default_rng(14)
    tight = 0
    for it in range(4000):
        scale = [1.0, 64.0, 4096.0][it % 3]
        lo = rng.uniform(-scale, scale, 3); hi = lo + rng.uniform(0, 1, 3) * (1.0 if it % 2 else scale)
        p = rng.uniform(-2 * scale, 2 * scale, 3)
        if it % 5 == 0:
            p = lo + rng.uniform(0, 1, 3) * (hi - lo)          # inside: distance 0
        exact = sum(max(Fr(float(l)) - Fr(float(x)), Fr(float(x)) - Fr(float(h)), 0) ** 2 for l, h, x in zip(lo, hi, p))
        got = hc.point_box_sqdist(np.concatenate([lo, hi]), p)
        assert Fr(got) <= exact, (lo, hi, p)
        if exact > 0 and float(exact) > 1e-3 * scale * scale:
            assert got >= float(exact) * (1 - 1e-3)             # and not uselessly loose
            tight += 1
    assert tight > 1000


@pytest.mark.parametrize("name", ["tetrahedron", "cube", "coverrear", "elephant"])
def test_lbvh_topology_and_boxes(meshes, name):
    v, f = meshes[name]
    t = hc.HostTree(soup_of(v, f))
    nodes, tris, bbox = t.export()
    assert hc.validate(nodes, tris, t.n) == 0
    np.testing.assert_array_equal(bbox, np.concatenate([v.min(0), v.max(0)]))
    ids = tris.view(np.uint32).reshape(-1, 24)[:, 18]
    assert sorted(ids.tolist()) == list(range(t.n))


@pytest.mark.parametrize("name", ["cube", "coverrear", "elephant"])
def test_queries_match_oracle(meshes, name):
    v, f = meshes[name]
    soup = soup_of(v, f)
    dev, ref = hc.HostTree(soup), orc.OracleTree(v, f)
    rng = np.random.default_rng(23)
    bb = ref.bbox()
    n = 1500
    q = rng.uniform(bb[:3], bb[3:], size=(n, 2, 3)).reshape(n, 6)
    for kind in (SEG, RAY):
        want = ref.count(q, kind)
        np.testing.assert_array_equal(dev.intersect(q, kind, True), want)
        np.testing.assert_array_equal(dev.intersect(q, kind, False), (want > 0).astype(np.uint32))
    # first_intersection: rays between points of the bounding sphere (aabb_test_ray_intersection.cpp:87-125)
    c, r = 0.5 * (bb[:3] + bb[3:]), 0.5 * (bb[3:] - bb[:3]).max()
    u = rng.normal(size=(n, 2, 3)); u /= np.linalg.norm(u, axis=2, keepdims=True)
    rays = (c + r * u).reshape(-1, 6)
    a, b = dev.first_intersection(rays), ref.first_intersection(rays)
    assert (b["prim"] != 0xFFFFFFFF).sum() > n // 10
    ok = a["prim"] == b["prim"]
    assert np.all(ok | (a["tie"] == 1) | (a["near"] == 1))
    sure = a["near"] == 0                              # outside the near-tie band t is bit-equal, inside within 1e-12
    np.testing.assert_array_equal(a["t"][sure], b["t"][sure])
    np.testing.assert_allclose(a["t"], b["t"], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(a["kind"], b["kind"])
    np.testing.assert_array_equal(a["pts"][ok], b["pts"][ok])
    # closest point: same point and id unless an exact-distance tie
    ext = bb[3:] - bb[:3]
    pts = rng.uniform(bb[:3] - 0.5 * ext, bb[3:] + 0.5 * ext, size=(n, 3))
    a, b = dev.closest_point(pts), ref.closest_point(pts)
    same = a["prim"] == b["prim"]
    assert np.all(same | (a["tie"] == 1) | (a["near"] == 1))
    np.testing.assert_array_equal(a["cp"][same], b["cp"][same])
    np.testing.assert_array_equal(a["sqd"][same], b["sqd"][same])
    for i in np.nonzero(~same & (a["near"] == 0))[0]:
        assert orc.compare_distance(pts[i], a["cp"][i], b["cp"][i]) == 0
        np.testing.assert_array_equal(ref.candidate_closest(a["prim"][i], pts[i]), a["cp"][i])
    assert same.mean() > 0.2   # far outside points mostly project on vertices / edges: exact ties


@pytest.mark.parametrize("name,warps,exact_go", [("cube", 1, 64), ("coverrear", 2, 32), ("elephant", 3, 64), ("elephant", 1, 128)])
def test_pair_kernel_body_on_emulated_warps(meshes, name, warps, exact_go):
    """closest_pair.cuh (the default closest_point kernel) run on the CPU by 32 cooperatively scheduled fibers per warp: same records as the
    thread-per-query traversal (bit for bit, tie flags included) and as the oracle up to reported exact ties; ragged batch sizes, queries far
    outside the box (ties on vertices / edges), with and without a coherence order"""
    v, f = meshes[name]
    dev, ref = hc.HostTree(soup_of(v, f)), orc.OracleTree(v, f)
    rng = np.random.default_rng(5)
    bb = ref.bbox()
    ext = bb[3:] - bb[:3]
    for n, order in ((1, False), (31, False), (777, True), (2500, False)):
        pts = rng.uniform(bb[:3] - 0.5 * ext, bb[3:] + 0.5 * ext, size=(n, 3))
        perm = rng.permutation(n).astype(np.uint32) if order else None
        a, t = dev.closest_point_pair(pts, perm=perm, warps=warps, exact_go=exact_go), dev.closest_point(pts)
        for k in ("cp", "prim", "sqd", "tie"):
            np.testing.assert_array_equal(a[k], t[k])
        b = ref.closest_point(pts)
        same = a["prim"] == b["prim"]
        assert np.all(same | (a["tie"] == 1))
        np.testing.assert_array_equal(a["sqd"], b["sqd"])
        np.testing.assert_array_equal(a["cp"][same], b["cp"][same])
        assert a["leaves"] >= n and a["nodes"] >= n


def test_pair_kernel_body_degenerate_and_deep_trees():
    """zero-area triangles (the reference's projection may not be the closest point: the first-vertex candidate rule), many duplicates
    (index tie-break levels of the Karras tree) and a chain-shaped tree (Morton keys 2^k: one level per key bit)"""
    rng = np.random.default_rng(9)
    base = rng.uniform(-1, 1, (40, 9))
    base[::4, 6:9] = base[::4, 3:6]                     # segments
    base[1::8, 3:6] = base[1::8, 0:3]; base[1::8, 6:9] = base[1::8, 0:3]   # points
    soup = np.concatenate([base, np.repeat(base[:5], 60, axis=0)])          # 300 duplicates
    chain = np.zeros((63, 9))
    for k in range(63):                                 # centres at one cell of a 2^21 grid per key bit: a 63-level chain
        c = np.zeros(3); c[2 - k % 3] = float(2 ** (k // 3))
        chain[k] = np.tile(c, 3) + np.array([0, 0, 0, 0.25, 0, 0, 0, 0.25, 0])
    chain = np.concatenate([chain, np.tile(np.array([2.0 ** 21, 2.0 ** 21, 2.0 ** 21]), 3)[None, :] + np.array([[0, 0, 0, 0.25, 0, 0, 0, 0.25, 0.0]])])
    for tris, lo, hi in ((soup, -2.0, 2.0), (chain, -10.0, 300.0)):
        dev = hc.HostTree(tris)
        pts = rng.uniform(lo, hi, size=(600, 3))
        a, t = dev.closest_point_pair(pts, warps=2, exact_go=48), dev.closest_point(pts)
        for k in ("cp", "prim", "sqd", "tie"):
            np.testing.assert_array_equal(a[k], t[k])
        # brute force over all candidates (projection, or the first vertex when strictly nearer), exact comparison
        o = orc.OracleTree(tris.reshape(-1, 3), np.arange(3 * len(tris), dtype=np.uint32).reshape(-1, 3))
        np.testing.assert_array_equal(a["sqd"], o.closest_point(pts)["sqd"])


def test_vertex_and_edge_ties_are_flagged(meshes):
    v, f = meshes["cube"]
    dev = hc.HostTree(soup_of(v, f))
    bb = np.concatenate([v.min(0), v.max(0)])
    far_corner = bb[3:] + 1.0                        # closest point is a cube vertex shared by several triangles
    r = dev.closest_point(far_corner[None, :])
    assert r["tie"][0] == 1 and tuple(r["cp"][0]) == tuple(bb[3:])
    ray = np.concatenate([bb[3:] + 1.0, bb[3:]])[None, :]   # ray through the same vertex: equal t on several triangles
    h = dev.first_intersection(ray)
    assert h["prim"][0] != 0xFFFFFFFF and h["tie"][0] == 1


def test_empty_and_singleton():
    e = hc.HostTree(np.zeros((0, 9)))
    q = np.array([[0.1, 0.1, -1.0, 0.1, 0.1, 1.0]])
    assert e.intersect(q, SEG, True)[0] == 0 and e.first_intersection(q)["prim"][0] == 0xFFFFFFFF
    s = hc.HostTree(np.array([[0, 0, 0, 1.0, 0, 0, 0, 1.0, 0]]), ids=[7])
    assert s.intersect(q, SEG, True)[0] == 1 and s.intersect(q, RAY, False)[0] == 1
    fi = s.first_intersection(q)
    assert fi["prim"][0] == 7 and tuple(fi["pts"][0][:3]) == (0.1, 0.1, 0.0) and fi["t"][0] == 0.5
    cp = s.closest_point(np.array([[0.25, 0.25, 5.0]]))
    assert tuple(cp["cp"][0]) == (0.25, 0.25, 0.0) and cp["prim"][0] == 7 and cp["sqd"][0] == 25.0


def _seg_dist(p, a, b):
    e = b - a
    ee = e @ e
    if ee == 0:
        return float(np.sqrt((p - a) @ (p - a)))
    t = min(max(((p - a) @ e) / ee, 0), 1)
    x = a + t * e
    return float(np.sqrt((p - x) @ (p - x)))


def _exact_point_triangle_distance(tri, q):
    """distance from q to the triangle in extended precision (Ericson, Real-Time Collision Detection 5.1.5); a triangle with a zero-length
    edge (vertices that coincide after rounding at extreme scale / offset ratios) is the union of its edges"""
    L = np.longdouble
    a, b, c, p = (np.array(x, dtype=L) for x in (tri[0:3], tri[3:6], tri[6:9], q))
    if not (np.any(a != b) and np.any(a != c) and np.any(b != c)):
        return min(_seg_dist(p, a, b), _seg_dist(p, a, c), _seg_dist(p, b, c))
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        return float(np.sqrt((p - a) @ (p - a)))
    bp = p - b
    d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3:
        return float(np.sqrt(bp @ bp))
    vc = d1 * d4 - d3 * d2
    if vc <= 0 and d1 >= 0 and d3 <= 0:
        x = a + (d1 / (d1 - d3)) * ab
        return float(np.sqrt((p - x) @ (p - x)))
    cp = p - c
    d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6:
        return float(np.sqrt(cp @ cp))
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0:
        x = a + (d2 / (d2 - d6)) * ac
        return float(np.sqrt((p - x) @ (p - x)))
    va = d3 * d6 - d5 * d4
    if va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
        x = b + ((d4 - d3) / ((d4 - d3) + (d5 - d6))) * (c - b)
        return float(np.sqrt((p - x) @ (p - x)))
    den = va + vb + vc
    if den == 0:                                        # degenerate triangle: the nearest of its three segments' end cases above
        return float(min(np.sqrt((p - v) @ (p - v)) for v in (a, b, c)))
    x = a + ab * (vb / den) + ac * (vc / den)
    return float(np.sqrt((p - x) @ (p - x)))


def test_warp_kernel_binary32_bounds_are_rigorous():
    """closest_warp.cuh prunes with binary32 LOWER bounds (stage 1: centroid direction, stage 2: approximate closest point) and shrinks a
    query's ball with a binary32 UPPER bound; every one must hold against the distance from the true (binary64) query to the true
    triangle, whatever the scale, the offset from the origin (binary32 copies lose everything) or the shape (needles, degenerate)"""
    rng = np.random.default_rng(7)
    n_upper = 0
    for trial in range(30000):
        scale = 10.0 ** (rng.uniform(-30, 30) if trial % 5 == 4 else rng.uniform(-6, 4))   # (every fifth pair: the edges of the binary32 range)
        offset = rng.choice([0.0, 1.0, 1e3, 1e7]) * rng.standard_normal(3)
        tri = rng.standard_normal(9) * scale
        kind = trial % 6
        if kind == 1:                                    # needle
            tri[6:9] = tri[0:3] + (tri[3:6] - tri[0:3]) * rng.uniform(0, 1) + 1e-9 * scale * rng.standard_normal(3)
        elif kind == 2:                                  # exactly degenerate: a segment / a point
            tri[6:9] = tri[3:6] if trial % 12 == 2 else tri[0:3]
        tri = tri + np.tile(offset, 3)
        dist = 10.0 ** rng.uniform(-3, 3) * scale
        q = tri[0:3] + rng.standard_normal(3) * dist
        if kind == 3:                                    # query on the triangle
            w = rng.dirichlet([1, 1, 1])
            q = w[0] * tri[0:3] + w[1] * tri[3:6] + w[2] * tri[6:9]
        mag = float(max(np.abs(tri).max(), np.abs(q).max()))
        lo1, lo2, up, e2, dslack = hc.warp_bounds(tri, q, mag)
        d = _exact_point_triangle_distance(tri, q)
        slack = 1e-9 * d + 1e-300
        assert lo1 <= d + slack and lo2 <= d + slack, (trial, tri, q, lo1, lo2, d)
        if not np.isnan(up):
            n_upper += 1
            assert up >= d - slack, (trial, tri, q, up, d)
            m_rel = float(np.abs(tri.reshape(3, 3) - q).max())
            assert up <= 1.001 * d + 4 * float(e2) + 2 * float(dslack) + 1e-3 * m_rel     # and it is useful (any point of the triangle is valid: ill-conditioned barycentrics only loosen it)
    assert n_upper > 10000                                # well-conditioned triangles do get an upper bound


def test_fused_bottom_up_build_equals_the_classic_build(meshes):
    """build.cuh::build_local_climb (topology + boxes of a block of 256 leaves in one bottom-up pass, then Karras' searches for the nodes that span
    blocks and the frontier climbs) produces, word for word, the nodes of hierarchy_kernel + refit: reference meshes, terrain, random soup, equal and
    duplicated primitives (index tie-break), a chain-shaped tree, and sizes around the block size"""
    from cgal_b200 import workloads as W
    rng = np.random.default_rng(5)
    soup = W.triangle_soup(40_000)
    cases = {name: meshes[name][0][meshes[name][1].astype(np.int64)].reshape(-1, 9) for name in ("cube", "tetrahedron", "elephant", "finger")}
    v, f = W.terrain(256, 128)
    cases["terrain"] = v[f.astype(np.int64)].reshape(-1, 9)
    cases["soup"] = soup
    cases["all equal"] = np.repeat(rng.uniform(0, 1, (1, 9)), 3000, 0)
    cases["duplicates"] = np.concatenate([soup[:5000], np.repeat(soup[:100], 50, 0)])
    k = np.arange(2000)
    x = 2.0 ** (-k / 40.0)
    c = np.stack([x, x, x], 1)
    cases["chain"] = np.concatenate([c, c + 1e-9 * x[:, None], c + 2e-9 * x[:, None]], 1)
    for n in (2, 3, 255, 256, 257, 511, 512, 513, 1025):
        cases["soup[:%d]" % n] = soup[:n]
    spanning = 0
    for name, tri in cases.items():
        diff, entries, rest, overflow = hc.HostTree(tri).fused_build_diff()
        assert (diff, overflow) == (0, 0), (name, diff, entries, rest, overflow)
        assert entries == (rest + 1 if rest else 0), (name, entries, rest)      # every spanning node is reached by exactly two parked children, the root included
        spanning += rest
    assert spanning > 1000
    # a frontier list that is too small is reported (the build then takes the classic path), never silently wrong
    assert hc.HostTree(cases["terrain"]).fused_build_diff(frontier_per_block=1)[3] == 1
