"""-m gpu: parity of the CUDA path (through the C ABI) with the CPU oracle.  Bit-exact for booleans, counts, ids
(except reported exact ties), ray parameters, hit points, closest points and squared distances."""
import numpy as np
import pytest

import hc
import orc
from gpu_util import check_closest_point, check_first_intersection, device_to_host

pytestmark = pytest.mark.gpu
SEG, RAY = 0, 1


@pytest.fixture(scope="module")
def B():
    import cgal_b200
    return cgal_b200


def soup_of(v, f):
    return np.ascontiguousarray(v[f.astype(np.int64)].reshape(-1, 9))


def test_kats_through_c_abi(B, meshes):
    # aabb_correctness_triangle_test.cpp:51-108
    v, f = meshes["tetrahedron"]
    t = B.AABB_tree(v, f)
    pq = np.array([[-0.25, 0.251, 0.255, 0.25, 0.253, 0.256]])
    assert t.do_intersect(pq, SEG)[0] == 1 and t.number_of_intersected_primitives(pq, SEG)[0] == 1
    assert tuple(t.closest_point(np.array([[0.0, 0.0, 3.0]]))[0]) == (0.0, 0.0, 1.0)
    # aabb_test_triangle_3.cpp:27-46
    a, b, c, d = (1.0, 0, 0), (0, 1.0, 0), (0, 0, 1.0), (0, 0, 0)
    t = B.AABB_tree(triangles=np.array([a + b + c, a + b + d, a + d + c], dtype=np.float64))
    assert t.number_of_intersected_primitives(np.array([a + b], dtype=np.float64), RAY)[0] == 3
    r = t.closest_point_and_primitive(np.array([[3.0, 2.0, 2.0]]))
    assert tuple(r["point"][0]) == a and r["sqdist"][0] == 12.0


def test_empty_singleton_and_state_machine(B):
    # aabb_test_empty_tree.cpp:38-53, aabb_test_singleton_tree.cpp:44-67
    q = np.array([[0.1, 0.1, -1.0, 0.1, 0.1, 1.0]])
    e = B.AABB_tree()
    assert e.size() == 0 and e.empty()
    assert e.do_intersect(q, SEG)[0] == 0 and e.number_of_intersected_primitives(q, RAY)[0] == 0
    assert e.first_intersection(q)["prim"][0] == B.NO_PRIMITIVE
    assert np.all(np.isinf(e.bbox()))
    with pytest.raises(B.B200AABBError) as ei:
        e.closest_point(np.zeros((1, 3)))
    assert ei.value.code == -4
    s = B.AABB_tree(triangles=np.array([[0, 0, 0, 1.0, 0, 0, 0, 1.0, 0]]), ids=[7])
    assert s.size() == 1
    assert s.do_intersect(q, SEG)[0] == 1 and s.number_of_intersected_primitives(q, RAY)[0] == 1
    fi = s.first_intersection(q)
    assert fi["prim"][0] == 7 and tuple(fi["point"][0]) == (0.1, 0.1, 0.0) and fi["t"][0] == 0.5 and fi["flags"][0] & B.FLAG_HIT
    cp = s.closest_point_and_primitive(np.array([[0.25, 0.25, 5.0]]))
    assert tuple(cp["point"][0]) == (0.25, 0.25, 0.0) and cp["prim"][0] == 7 and cp["sqdist"][0] == 25.0
    np.testing.assert_array_equal(s.bbox(), [0, 0, 0, 1, 1, 0])
    s.clear()
    assert s.empty() and s.do_intersect(q, SEG)[0] == 0
    s.insert(triangles=np.array([[0, 0, 0, 1.0, 0, 0, 0, 1.0, 0], [0, 0, 1, 1.0, 0, 1, 0, 1.0, 1]]))   # insert after clear, lazy rebuild
    assert s.size() == 2 and s.number_of_intersected_primitives(q, RAY)[0] == 2
    assert s.number_of_intersected_primitives(q, SEG)[0] == 2          # the target endpoint lies ON the z = 1 triangle: closed sets
    assert s.number_of_intersected_primitives(np.array([[0.1, 0.1, -1.0, 0.1, 0.1, 0.5]]), SEG)[0] == 1
    assert s.number_of_intersected_primitives(np.array([[0.1, 0.1, -1.0, 0.1, 0.1, 2.0]]), SEG)[0] == 2


@pytest.mark.parametrize("name", ["cube", "coverrear", "elephant", "finger"])
def test_gpu_built_tree_equals_host_emulation(B, meshes, name):
    """radix sort + Karras + refit on the GPU reproduce the step-by-step host emulation of the same algorithm bit for bit"""
    v, f = meshes[name]
    t = B.AABB_tree(v, f)
    t.build()
    (pn, bn), (pt, bt), (pb, bb), (pf, bf) = t.export_layout()
    n = len(f)
    assert bn == (n - 1) * 64 and bt == n * 96 and bb == 48 and bf == n * 64
    g_nodes, g_tris = device_to_host(pn, bn), device_to_host(pt, bt)
    assert hc.validate(g_nodes, g_tris, n) == 0
    h_nodes, h_tris, h_bbox = hc.HostTree(soup_of(v, f)).export()
    np.testing.assert_array_equal(g_tris, h_tris)
    want_f = g_tris.view(np.float64).reshape(-1, 12)[:, :9].astype(np.float32)     # binary32 copies: round to nearest
    np.testing.assert_array_equal(device_to_host(pf, bf).view(np.float32).reshape(-1, 16)[:, :9], want_f)
    gw, hw = g_nodes.view(np.uint32).reshape(-1, 16), h_nodes.view(np.uint32).reshape(-1, 16)
    cols = list(range(0, 15))                   # boxes, children, parent (not the refit arrival counter)
    np.testing.assert_array_equal(gw[:, cols], hw[:, cols])
    np.testing.assert_array_equal(t.bbox(), h_bbox)
    np.testing.assert_array_equal(device_to_host(pb, 48).view(np.float64), h_bbox)


def test_large_random_soup_tree_is_valid(B):
    from cgal_b200 import workloads as W
    n = 1_000_003           # ragged last sort tile, many Morton prefixes
    soup = W.triangle_soup(n)
    t = B.AABB_tree(triangles=soup)
    t.build()
    (pn, bn), (pt, bt) = t.export_layout()[:2]
    g_nodes, g_tris = device_to_host(pn, bn), device_to_host(pt, bt)
    assert hc.validate(g_nodes, g_tris, n) == 0
    ids = g_tris.view(np.uint32).reshape(-1, 24)[:, 18]
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))
    np.testing.assert_array_equal(t.bbox(), np.concatenate([soup.reshape(-1, 3).min(0), soup.reshape(-1, 3).max(0)]))
    # duplicates: 50k copies of 4 triangles -> equal Morton keys resolved by index
    dup = np.tile(soup[:4], (50_000, 1))
    t2 = B.AABB_tree(triangles=dup)
    t2.build()
    (pn, bn), (pt, bt) = t2.export_layout()[:2]
    assert hc.validate(device_to_host(pn, bn), device_to_host(pt, bt), len(dup)) == 0


def test_degenerate_tree_shapes(B):
    """a deep, unbalanced LBVH (geometrically growing coordinates), built with the refit's frontier list disabled so that every
    hand-over takes the in-place climb (test hook, option 101) and with the list enabled: same tree; results equal the oracle"""
    n = 20000
    s = 1.0005 ** np.arange(n)
    tri = np.stack([s, s, s, s * 1.0001, s, s, s, s * 1.0001, s], 1)
    g, o = B.AABB_tree(triangles=tri), orc.OracleTree(soup=tri)
    g.set_option(101, 0)
    g.build()
    (pn, bn), (pt, bt) = g.export_layout()[:2]
    nodes0 = device_to_host(pn, bn).view(np.uint32).reshape(-1, 16)[:, :15].copy()
    assert hc.validate(device_to_host(pn, bn), device_to_host(pt, bt), n) == 0
    g.set_option(101, 16)
    g.build()
    (pn, bn), (pt, bt) = g.export_layout()[:2]
    np.testing.assert_array_equal(device_to_host(pn, bn).view(np.uint32).reshape(-1, 16)[:, :15], nodes0)
    pts = np.stack([s[::7] * 1.00005, s[::7], s[::7] * 0.9999], 1)
    check_closest_point(g.closest_point_and_primitive(pts), o.closest_point(pts), o, pts)
    segs = np.concatenate([pts, pts * 1.01], 1)
    np.testing.assert_array_equal(g.number_of_intersected_primitives(segs, SEG), o.count(segs, SEG))


def test_packed_and_full_key_builds(B, meshes):
    """the build sorts PACKED words (top key bits | primitive index, build.cuh::pack_sort_word) and falls back to full 63-bit keys + values when
    many primitives share their kept key bits: same tree on a mesh without such pairs; a crowded scene takes the fall-back; both equal
    the host emulation"""
    v, f = meshes["elephant"]
    t = B.AABB_tree(v, f)
    t.build()
    (pn, bn) = t.export_layout()[0]
    packed = device_to_host(pn, bn).view(np.uint32).reshape(-1, 16)[:, :15].copy()
    t.set_option(102, 1)             # test hook: full keys
    t.build()
    (pn, bn) = t.export_layout()[0]
    np.testing.assert_array_equal(device_to_host(pn, bn).view(np.uint32).reshape(-1, 16)[:, :15], packed)
    # 60,000 triangles crowded into 1e-9 of the scene's extent + one far away: most neighbours agree in the kept key bits
    rng = np.random.default_rng(3)
    tri = rng.random((60000, 9)) * 1e-9
    tri = np.concatenate([tri, [[1.0, 1.0, 1.0, 1.0, 1.1, 1.0, 1.1, 1.0, 1.0]]])
    g = B.AABB_tree(triangles=tri)
    g.build()
    (pn, bn), (pt, bt) = g.export_layout()[:2]
    assert hc.validate(device_to_host(pn, bn), device_to_host(pt, bt), len(tri)) == 0
    h_nodes, h_tris, _ = hc.HostTree(tri).export()
    np.testing.assert_array_equal(device_to_host(pt, bt), h_tris)
    np.testing.assert_array_equal(device_to_host(pn, bn).view(np.uint32).reshape(-1, 16)[:, :15], h_nodes.view(np.uint32).reshape(-1, 16)[:, :15])
    o = orc.OracleTree(soup=tri)
    pts = rng.random((3000, 3)) * 2e-9
    check_closest_point(g.closest_point_and_primitive(pts), o.closest_point(pts), o, pts)


def test_chain_shaped_tree_depth(B):
    """the deepest trees the builder can produce: Morton keys 2^k give one level per key bit (a 63-level chain), 4,096 coincident primitives
    add 12 index levels below it; rays, segments and points equal the oracle and no traversal reports an overflow (closest_pair.cuh: depth
    <= 94 by construction, stacks hold 96)"""
    chain = np.zeros((63, 9))
    for k in range(63):
        c = np.zeros(3); c[2 - k % 3] = float(2 ** (k // 3))
        chain[k] = np.tile(c, 3) + np.array([0, 0, 0, 0.25, 0, 0, 0, 0.25, 0])
    top = np.tile(np.array([2.0 ** 21] * 3), 3) + np.array([0, 0, 0, 0.25, 0, 0, 0, 0.25, 0.0])
    tri = np.concatenate([chain, top[None, :], np.tile(chain[:1], (4096, 1))])
    g, o = B.AABB_tree(triangles=tri), orc.OracleTree(soup=tri)
    rng = np.random.default_rng(4)
    pts = np.concatenate([rng.uniform(-2, 40, (3000, 3)), rng.uniform(-2, 2.0 ** 21, (1000, 3))])
    for kernel in (1, 2, 3, 0):
        g.set_option(B.OPT_CLOSEST_KERNEL, kernel)
        check_closest_point(g.closest_point_and_primitive(pts), o.closest_point(pts), o, pts)
        g.stats()                      # raises if the batch reported an incomplete traversal
    segs = np.concatenate([pts, pts[::-1]], 1)
    np.testing.assert_array_equal(g.number_of_intersected_primitives(segs, SEG), o.count(segs, SEG))
    np.testing.assert_array_equal(g.number_of_intersected_primitives(segs, RAY), o.count(segs, RAY))
    rays = np.concatenate([pts + [0, 0, 5.0], pts - [0, 0, 5.0]], 1)
    check_first_intersection(g.first_intersection(rays), o.first_intersection(rays), o, rays, max_mismatch_frac=1.0)
    g.stats()


def test_config1_elephant_parity(B, meshes):
    """BASELINE.json configs[0]: elephant.off, random rays + closest-point queries, CPU oracle on all queries"""
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    bb = g.bbox()
    np.testing.assert_array_equal(bb, o.bbox())
    n = 200_000
    rays = W.sphere_rays(bb, n)
    gr, rr = g.first_intersection(rays), o.first_intersection(rays)
    assert (rr["prim"] != B.NO_PRIMITIVE).mean() > 0.05
    mism, ties = check_first_intersection(gr, rr, o, rays)
    assert mism <= ties
    pts = W.bbox_points(bb, n)
    gc, rc = g.closest_point_and_primitive(pts), o.closest_point(pts)
    check_closest_point(gc, rc, o, pts)
    np.testing.assert_array_equal(g.squared_distance(pts[:1000]), gc["sqdist"][:1000])
    segs = W.bbox_segments(bb, n)
    for kind, q in ((SEG, segs), (RAY, rays), (RAY, segs)):
        want = o.count(q, kind)
        np.testing.assert_array_equal(g.number_of_intersected_primitives(q, kind), want)
        np.testing.assert_array_equal(g.do_intersect(q, kind), (want > 0).astype(np.uint8))


def test_all_and_any_intersected_primitives(B, meshes):
    """all_intersected_primitives / any_intersected_primitive (AABB_tree.h:934-956): the id SET of every query equals the oracle's
    (the order within a query is the traversal order, unspecified in the reference too); `any` returns a member of the set"""
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    bb = g.bbox()
    for kind, q in ((SEG, W.bbox_segments(bb, 60_000)), (RAY, W.sphere_rays(bb, 60_000))):
        off, ids = g.all_intersected_primitives(q, kind)
        roff, rids = o.all_intersected_primitives(q, kind)
        np.testing.assert_array_equal(off, roff)
        assert off[-1] > 10_000
        owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
        key = lambda own, i: np.sort(own.astype(np.uint64) << np.uint64(32) | i.astype(np.uint64))   # noqa: E731
        np.testing.assert_array_equal(key(owner, ids), key(owner, rids))                   # same (query, id) pairs
        a = g.any_intersected_primitive(q, kind)
        cnt = np.diff(off)
        np.testing.assert_array_equal(a == B.NO_PRIMITIVE, cnt == 0)
        pairs = set(zip(owner.tolist(), ids.tolist()))
        assert all((i, int(a[i])) in pairs for i in np.nonzero(cnt)[0][:5000])
    # wrong offsets are refused, and nothing is written outside a query's own range
    q = W.bbox_segments(bb, 1000)
    cnt = g.number_of_intersected_primitives(q, SEG)
    off = np.zeros(len(q) + 1, np.uint64)
    np.cumsum(cnt, out=off[1:])
    bad = off.copy()
    k = int(np.nonzero(cnt)[0][0])
    bad[k + 1:] -= np.uint64(1)                      # one id short for query k
    ids = np.full(int(off[-1]) + 8, 0xDEADBEEF, np.uint32)
    rc = g._lib.b200aabb_all_intersected_primitives(g._h, q.ctypes.data, len(q), SEG, 0, bad.ctypes.data, ids.ctypes.data)
    assert rc == -1 and np.all(ids[int(bad[-1]):] == 0xDEADBEEF)
    # empty tree
    e = B.AABB_tree()
    assert e.any_intersected_primitive(q[:3], SEG).tolist() == [B.NO_PRIMITIVE] * 3
    eo, ei = e.all_intersected_primitives(q[:3], SEG)
    assert eo.tolist() == [0, 0, 0, 0] and len(ei) == 0


def test_all_and_any_intersections_of_rays(B, meshes):
    """all_intersections / any_intersection (Ray_3), AABB_tree.h:958-985: the set of (ray, primitive, object) records equals the oracle's,
    objects bit for bit; first_intersection is the nearest member of a ray's set (aabb_test_ray_intersection.cpp:38-65)"""
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    rays = W.sphere_rays(g.bbox(), 60_000)
    # axis-parallel rays through vertices and along edges of a grid mesh: coplanar hits (Segment_3 objects), exact stage
    gv, gf = W.terrain(24, 24, amplitude=0.0)
    flat, oflat = B.AABB_tree(gv, gf), orc.OracleTree(gv, gf)
    xs = np.arange(0.0, 25.0)
    inplane = np.stack([np.full(25, -1.0), xs, np.zeros(25), np.full(25, 0.0), xs, np.zeros(25)], 1)          # along grid lines, in the plane
    through = np.stack([xs, xs[::-1], np.full(25, 5.0), xs, xs[::-1], np.full(25, 4.0)], 1)                       # through vertices
    for tree, orac, q in ((g, o, rays), (flat, oflat, np.concatenate([inplane, through]))):
        off, ids, kinds, pts = tree.all_intersections(q)
        roff, rids, rkinds, rpts = orac.all_intersections(q)
        np.testing.assert_array_equal(off, roff)
        owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
        order = np.lexsort((ids, owner)); rorder = np.lexsort((rids, owner))
        np.testing.assert_array_equal(ids[order], rids[rorder])
        np.testing.assert_array_equal(kinds[order], rkinds[rorder])
        np.testing.assert_array_equal(pts[order], rpts[rorder])                                # objects bit-equal
        a = tree.any_intersection(q)
        cnt = np.diff(off)
        np.testing.assert_array_equal(a["prim"] == B.NO_PRIMITIVE, cnt == 0)
        np.testing.assert_array_equal((a["flags"] & B.FLAG_HIT) != 0, cnt > 0)
        fi = tree.first_intersection(q)
        lookup = {(int(w), int(i)): k for k, (w, i) in enumerate(zip(owner, ids))}
        for i in np.nonzero(cnt)[0][:4000]:
            k = lookup[(int(i), int(a["prim"][i]))]                                              # a member of the ray's set, with its object
            assert np.array_equal(a["point"][i], pts[k, :3]) and bool(a["flags"][i] & B.FLAG_SEGMENT) == (kinds[k] == 2)
            assert (int(i), int(fi["prim"][i])) in lookup
    assert (kinds == 2).any()                                                                    # the grid-line rays produced Segment_3 objects
    e = B.AABB_tree()
    assert e.any_intersection(rays[:3])["prim"].tolist() == [B.NO_PRIMITIVE] * 3 and len(e.all_intersections(rays[:3])[1]) == 0


def test_config2_icosphere_distance_field_parity(B):
    from cgal_b200 import workloads as W
    v, f = W.icosphere(6)       # 81,920 triangles (config 2 uses level 8; same generator)
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    pts = W.grid_points(64)     # 262,144 grid points over [-1.5,1.5]^3
    gc, rc = g.closest_point_and_primitive(pts), o.closest_point(pts)
    mism, ties = check_closest_point(gc, rc, o, pts)
    r = np.sqrt(gc["sqdist"])
    assert np.all(np.abs(np.linalg.norm(gc["point"], axis=1) - 1) < 1e-3)      # closest points lie on the unit icosphere
    assert np.all(np.abs(r - np.abs(np.linalg.norm(pts, axis=1) - 1)) < 2e-3)  # unsigned distance field of a sphere


@pytest.mark.parametrize("exact_go,sort", [(128, 1), (1, 1), (160, 0), (16, 1)])
def test_closest_point_kernels_agree(B, exact_go, sort):
    """the warp-cooperative work-pool kernel (closest_warp.cuh) and the thread-per-query kernel (bvh.cuh) return identical
    records, whatever the pool scheduling parameters and the query order; both equal the oracle"""
    from cgal_b200 import workloads as W
    v, f = W.terrain(256, 128)      # 65,536 triangles
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    bb = g.bbox()
    pts = np.concatenate([W.bbox_points(bb, 150_000, margin=0.2), v[::7] + 1e-9, v[5::11],            # near / on vertices: ties
                          0.5 * (v[f[::13, 0]] + v[f[::13, 1]])])                                     # on edges
    g.set_option(B.OPT_SORT_QUERIES, sort)
    g.set_option(B.OPT_CLOSEST_KERNEL, 0)
    ref = g.closest_point_and_primitive(pts)
    for kernel in (1, 2, 3):         # the work-pool kernel, the lane-pair kernel (closest_pair.cuh) and the two-pass kernels (binary32 filter + exact resolve)
        g.set_option(B.OPT_CLOSEST_KERNEL, kernel)
        g.set_option(B.OPT_WARP_EXACT_GO, min(exact_go, 128) if kernel == 2 else exact_go)
        got = g.closest_point_and_primitive(pts)
        st = g.stats()
        assert st["nodes_visited"] > 0 and st["leaves_tested"] > 0
        for k in ("point", "prim", "sqdist"):
            np.testing.assert_array_equal(got[k], ref[k], err_msg="%s kernel %d" % (k, kernel))
        np.testing.assert_array_equal(got["flags"] & B.FLAG_TIE, ref["flags"] & B.FLAG_TIE)   # (FLAG_EXACT depends on which incumbents met)
        check_closest_point(got, o.closest_point(pts), o, pts)
    g.set_option(B.OPT_CLOSEST_KERNEL, 1)
    g.set_option(B.OPT_WARP_EXACT_GO, exact_go)
    # tiny batches and trees: fewer queries than slots, two-primitive tree
    few = g.closest_point_and_primitive(pts[:5])
    np.testing.assert_array_equal(few["prim"], ref["prim"][:5])
    t2 = B.AABB_tree(triangles=np.array([[0, 0, 0, 1.0, 0, 0, 0, 1.0, 0], [0, 0, 1, 1.0, 0, 1, 0, 1.0, 1]]))
    r2 = t2.closest_point_and_primitive(np.array([[0.2, 0.2, 0.4], [0.2, 0.2, 0.6], [0.2, 0.2, 0.5]]))
    assert list(r2["prim"]) == [0, 1, 0] and r2["flags"][2] & B.FLAG_TIE and not (r2["flags"][0] & B.FLAG_TIE)


@pytest.mark.parametrize("kernel", [1, 2, 3, 0])
def test_closest_point_adversarial_inputs(B, kernel):
    """inputs that stress the binary32 pruning aids of closest_point (they may only ever fail to prune): degenerate and needle
    triangles, thousands of coincident primitives (exact ties), coordinates far from the origin (binary32 copies useless),
    a tiny scene; both kernels must return the oracle's records"""
    from cgal_b200 import workloads as W
    rng = np.random.default_rng(23)

    def check(tri, pts, min_ties=0):
        g, o = B.AABB_tree(triangles=tri), orc.OracleTree(soup=tri)
        g.set_option(B.OPT_CLOSEST_KERNEL, kernel)
        got = g.closest_point_and_primitive(pts)
        mism, ties = check_closest_point(got, o.closest_point(pts), o, pts)
        assert ties >= min_ties
        return got

    base = W.triangle_soup(3000)
    # zero-area triangles (a point, a segment), needles, and well-shaped ones mixed
    deg = base.copy()
    deg[0::5, 3:6] = deg[0::5, 0:3]; deg[0::5, 6:9] = deg[0::5, 0:3]                   # points
    deg[1::5, 6:9] = 0.5 * (deg[1::5, 0:3] + deg[1::5, 3:6])                          # collinear
    deg[2::5, 6:9] = deg[2::5, 0:3] + 1e-13 * rng.standard_normal((len(deg[2::5]), 3))   # needles
    pts = rng.random((20000, 3))
    check(deg, pts)
    check(deg, np.concatenate([deg[0::5, 0:3], deg[1::5, 6:9]]))                      # queries ON the degenerate primitives
    # 4 triangles, 2500 copies each: every query ties 2500-fold; smallest id wins and the tie is reported
    dup = np.tile(base[:4], (2500, 1))
    got = check(dup, pts[:3000], min_ties=3000)
    assert np.all(got["prim"] < 4) and np.all(got["flags"] & B.FLAG_TIE)
    # far from the origin: |coordinates| ~ 1e7 with unit-size scene (binary32 ulp = 1) and a micro scene at 1e-6 scale
    far = base + 1e7
    check(far, pts[:5000] + 1e7)
    tiny = base * 1e-6
    check(tiny, pts[:5000] * 1e-6)
    # the edges of the binary32 range (squares under / overflow in binary32: the pre-tests must go quiet, not wrong)
    for scale in (1e-21, 1e-13, 1e13, 1e20, 1e30):
        check(base[:600] * scale, pts[:1500] * scale)
    # a large flat grid seen from far away: thousands of leaves within the rounding allowance of the same distance
    v, f = W.terrain(128, 128, amplitude=0.0)
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    g.set_option(B.OPT_CLOSEST_KERNEL, kernel)
    hi = np.stack([rng.random(2000) * 128, rng.random(2000) * 128, 1e4 + rng.random(2000)], 1)
    check_closest_point(g.closest_point_and_primitive(hi), o.closest_point(hi), o, hi)


def test_config3_terrain_rays_and_shadow_segments_parity(B):
    from cgal_b200 import workloads as W
    v, f = W.terrain(512, 256)  # 262,144 triangles (config 3 uses 4096 x 2048; same generator)
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    rays = W.pinhole_rays(512, 512, 512, 256)
    gr, rr = g.first_intersection(rays), o.first_intersection(rays)
    assert (rr["prim"] != B.NO_PRIMITIVE).mean() > 0.2
    check_first_intersection(gr, rr, o, rays)
    segs = W.shadow_segments(v, 200_000, length=512.0)
    want = o.do_intersect(segs, SEG)
    assert 0.02 < want.mean() < 0.98
    np.testing.assert_array_equal(g.do_intersect(segs, SEG), want)
    np.testing.assert_array_equal(g.number_of_intersected_primitives(segs, SEG), o.count(segs, SEG))
    # axis-parallel and grid-aligned queries: vertical rays through grid vertices / along grid lines (exact stage, ties)
    xs, ys = np.meshgrid(np.arange(0, 513, 8, dtype=np.float64), np.arange(0, 257, 8, dtype=np.float64))
    down = np.stack([xs.ravel(), ys.ravel(), np.full(xs.size, 100.0), xs.ravel(), ys.ravel(), np.full(xs.size, 99.0)], 1)
    gd, rd = g.first_intersection(down), o.first_intersection(down)
    # every ray passes exactly through a vertex shared by six triangles whose rounded t differ by an ulp: near ties by construction
    mism, ties = check_first_intersection(gd, rd, o, down, max_mismatch_frac=0.05)
    assert ties > 0 and g.stats()["exact_fallbacks"] > 0
    np.testing.assert_array_equal(g.number_of_intersected_primitives(down, RAY), o.count(down, RAY))


def test_config4_soup_counts_parity(B):
    from cgal_b200 import workloads as W
    soup = W.triangle_soup(200_000)
    g, o = B.AABB_tree(triangles=soup), orc.OracleTree(soup=soup)
    segs = W.short_segments(100_000)
    want = o.count(segs, SEG)
    assert want.max() >= 2
    np.testing.assert_array_equal(g.number_of_intersected_primitives(segs, SEG), want)
    np.testing.assert_array_equal(g.do_intersect(segs, SEG), (want > 0).astype(np.uint8))


def test_device_pointers_and_host_pointers_agree(B, meshes):
    import torch
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g = B.AABB_tree(v, f)
    rays = W.sphere_rays(g.bbox(), 5_000_000 // 100)
    h = g.first_intersection(rays)
    d = g.first_intersection(torch.from_numpy(rays).cuda())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d["prim"].cpu().numpy().view(np.uint32), h["prim"])
    np.testing.assert_array_equal(d["t"].cpu().numpy(), h["t"])
    np.testing.assert_array_equal(d["flags"].cpu().numpy(), h["flags"])
    # a HOST batch larger than one pipeline chunk (4 Mi queries) exercises the double-buffered copy pipeline
    n = (1 << 22) + 12345
    pts = W.bbox_points(g.bbox(), n)
    a = g.closest_point_and_primitive(pts)
    b = g.closest_point_and_primitive(torch.from_numpy(pts).cuda())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(b["prim"].cpu().numpy().view(np.uint32), a["prim"])
    np.testing.assert_array_equal(b["point"].cpu().numpy(), a["point"])


def test_full_size_properties_16m_terrain(B):
    """BASELINE.json configs[2] at full size: 16,777,216 triangles.  Size-independent properties instead of the oracle."""
    import torch
    from cgal_b200 import workloads as W
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.build()
    assert g.size() == 16_777_216
    np.testing.assert_array_equal(g.bbox(), np.concatenate([v.min(0), v.max(0)]))
    rays = W.pinhole_rays(8192, 8192, rows=(4096, 4096 + 128))        # 1 Mi primary rays from the image centre rows
    r = g.first_intersection(rays)
    hit = r["prim"] != B.NO_PRIMITIVE
    assert hit.mean() > 0.5
    np.testing.assert_array_equal(g.do_intersect(rays, RAY), hit.astype(np.uint8))       # any-hit agrees with first-hit
    assert np.all(r["t"][hit] > 0)
    tri = v[f[r["prim"][hit]].astype(np.int64)]                                            # hit point lies on the reported triangle
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    off = np.abs(np.einsum("ij,ij->i", r["point"][hit] - tri[:, 0], nrm)) / np.linalg.norm(nrm, axis=1)
    assert off.max() < 1e-9
    # a segment cut just short of the hit misses, one extending past it hits (first hit is the FIRST hit)
    s, d = rays[hit][:, :3], rays[hit][:, 3:] - rays[hit][:, :3]
    t = r["t"][hit][:, None]
    short = np.concatenate([s, s + d * t * (1 - 1e-6)], 1)
    longer = np.concatenate([s, s + d * t * (1 + 1e-6)], 1)
    assert g.do_intersect(short, SEG).sum() == 0
    # (a ray grazing a silhouette edge may slip past it once its direction is re-rounded: allow 1e-4 of the rays)
    assert g.do_intersect(longer, SEG).mean() > 1 - 1e-4
    # closest point: idempotent and never farther than a brute-force scan over a vertex sample
    pts = W.bbox_points(g.bbox(), 1 << 20)
    c = g.closest_point_and_primitive(pts)
    again = g.closest_point_and_primitive(c["point"])
    assert again["sqdist"].max() < 1e-18 * (4096.0 ** 2)
    sample = v[:: 4099]
    dv = ((pts[:2000, None, :] - sample[None, :, :]) ** 2).sum(-1).min(1)
    assert np.all(c["sqdist"][:2000] <= dv * (1 + 1e-12))


def test_first_intersection_with_skip(B, meshes):
    """first_intersection(ray, skip) (AABB_tree.h:307-328): per-ray skipped primitive (the face the ray leaves) and a mask over primitive ids,
    against the oracle's restatement; a one-primitive tree ignores skip, as the reference does"""
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    rays = W.sphere_rays(g.bbox(), 20000)
    first = g.first_intersection(rays)
    hit = first["prim"] != B.NO_PRIMITIVE
    # second hit along each ray = first hit when the first primitive is skipped
    got = g.first_intersection_skip(rays, skip_ids=first["prim"])
    ref = o.first_intersection_skip(rays, skip_ids=first["prim"])
    tie = (got["flags"] & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
    ok = got["prim"] == ref["prim"]
    assert np.all(ok | tie)
    np.testing.assert_array_equal(got["t"][ok], ref["t"][ok])
    assert np.all(got["prim"][hit] != first["prim"][hit]) and np.all(got["t"][hit & (got["prim"] != B.NO_PRIMITIVE)] >= first["t"][hit & (got["prim"] != B.NO_PRIMITIVE)])
    mask = (np.arange(len(f)) % 3) == 0
    got = g.first_intersection_skip(rays, skip_mask=mask)
    ref = o.first_intersection_skip(rays, skip_mask=mask)
    tie = (got["flags"] & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
    ok = got["prim"] == ref["prim"]
    assert np.all(ok | tie) and not np.any(mask[got["prim"][got["prim"] != B.NO_PRIMITIVE]])
    np.testing.assert_array_equal(got["t"][ok], ref["t"][ok])
    one = B.AABB_tree(triangles=np.array([[0, 0, 0, 1.0, 0, 0, 0, 1.0, 0]]), ids=[5])
    r = one.first_intersection_skip(np.array([[0.1, 0.1, -1.0, 0.1, 0.1, 1.0]]), skip_ids=[5])
    assert r["prim"][0] == 5


def test_max_distance_to_mesh(B, meshes):
    """PMP::max_distance_to_mesh_impl (distance.h:98-190) = sqrt of the largest squared distance of the samples; host and device pointers"""
    import torch
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    pts = W.bbox_points(g.bbox(), 50000, margin=0.3)
    sq = o.closest_point(pts)["sqd"]
    d, i = g.max_distance_to_mesh(pts)
    assert d == np.sqrt(sq.max()) and sq[i] == sq.max()
    d2, i2 = g.max_distance_to_mesh(torch.from_numpy(pts).cuda())
    assert (d2, i2) == (d, i)


def test_refit_keeps_results_of_a_fresh_build(B, meshes):
    """b200aabb_refit / _refit_mesh: the primitives move, the hierarchy stays; boxes are recomputed bottom-up (valid tree) and every query
    answers like a freshly built tree and like the oracle on the deformed mesh; twice in a row (the refit's arrival flags reset themselves)"""
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g = B.AABB_tree(v, f)
    g.build()
    rng = np.random.default_rng(8)
    for rnd in range(2):
        v2 = v + 0.02 * (v.max(0) - v.min(0)) * rng.standard_normal(v.shape) * (rnd + 1)
        if rnd == 0:
            g.refit(vertices=v2, faces=f)
        else:
            g.refit(triangles=soup_of(v2, f))
        (pn, bn), (pt, bt) = g.export_layout()[:2]
        assert hc.validate(device_to_host(pn, bn), device_to_host(pt, bt), len(f)) == 0
        o = orc.OracleTree(v2, f)
        assert np.array_equal(g.bbox(), o.bbox())
        bb = g.bbox()
        rays, pts, segs = W.sphere_rays(bb, 5000), W.bbox_points(bb, 5000), W.bbox_segments(bb, 5000)
        check_first_intersection(g.first_intersection(rays), o.first_intersection(rays), o, rays)
        check_closest_point(g.closest_point_and_primitive(pts), o.closest_point(pts), o, pts)
        np.testing.assert_array_equal(g.number_of_intersected_primitives(segs, SEG), o.count(segs, SEG))
    with pytest.raises(B.B200AABBError):
        g.refit(triangles=soup_of(v, f)[:-1])      # the number of primitives must not change


def test_async_host_batches(B, meshes):
    """B200AABB_OPT_ASYNC_HOST: HOST batches return once enqueued and overlap across calls; after b200aabb_synchronize the results equal the
    synchronous ones"""
    import torch
    from cgal_b200 import workloads as W
    v, f = W.terrain(256, 128)
    g = B.AABB_tree(v, f)
    bb = g.bbox()
    n = 300_000
    rays = torch.from_numpy(W.pinhole_rays(600, 500, 256, 128)).pin_memory().numpy()
    pts = torch.from_numpy(W.bbox_points(bb, n)).pin_memory().numpy()
    want_r = g.first_intersection(rays)
    want_p = g.closest_point_and_primitive(pts)
    g.set_option(B.OPT_ASYNC_HOST, 1)
    outs = []
    for _ in range(3):
        outs.append((g.first_intersection(rays), g.closest_point_and_primitive(pts)))
    g.synchronize()
    g.set_option(B.OPT_ASYNC_HOST, 0)
    for r, p in outs:
        for k in want_r:
            np.testing.assert_array_equal(r[k], want_r[k])
        for k in want_p:
            np.testing.assert_array_equal(p[k], want_p[k])


@pytest.mark.parametrize("name", ["elephant", "cube"])
def test_line_and_plane_queries(B, meshes, name):
    """Line_3 (point + direction vector) and Plane_3 (a, b, c, d) queries — do_intersect, number_of_intersected_primitives,
    any_ / all_intersected_primitives — against the oracle's exact evaluation of Line_3_Triangle_3_do_intersect.h:27-79 and
    Plane_3_Triangle_3_do_intersect.h:27-51; random queries plus lines through vertices / along edges / in facet planes and planes
    through vertices / facets (tests/test_line_plane_queries.py builds them); host and device pointers"""
    import torch
    from test_line_plane_queries import _mesh_queries
    v, f = meshes[name]
    rng = np.random.default_rng(21)
    lines, planes = _mesh_queries(rng, v, f, 20_000 if name == "elephant" else 2_000)
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    for kind, q in ((B.LINE, lines), (B.PLANE, planes)):
        want = o.count(q, kind)
        cnt = g.number_of_intersected_primitives(q, kind)
        np.testing.assert_array_equal(cnt, want)
        np.testing.assert_array_equal(g.do_intersect(q, kind) != 0, want > 0)
        assert want.max() > 2 and (want == 0).any()
        off, ids = g.all_intersected_primitives(q, kind)
        roff, rids = o.all_intersected_primitives(q, kind)
        np.testing.assert_array_equal(off, roff)
        owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
        key = lambda own, i: np.sort(own.astype(np.uint64) << np.uint64(32) | i.astype(np.uint64))   # noqa: E731
        np.testing.assert_array_equal(key(owner, ids), key(owner, rids))
        a = g.any_intersected_primitive(q, kind)
        np.testing.assert_array_equal(a == B.NO_PRIMITIVE, want == 0)
        pairs = set(zip(owner.tolist(), ids.tolist()))
        assert all((i, int(a[i])) in pairs for i in np.nonzero(want)[0][:3000])
        qd = torch.from_numpy(q).cuda()
        np.testing.assert_array_equal(g.number_of_intersected_primitives(qd, kind).cpu().numpy(), want)
    with pytest.raises(B.B200AABBError):
        g.do_intersect(lines[:4], 7)


def test_fused_and_classic_build_paths_agree(B, meshes):
    """the fused bottom-up build (build.cuh::build_local_kernel + searches for the spanning nodes) and the classic path (Karras' searches for every
    node + refit; test hook 103, also taken when the fused path's frontier list fills up: hook 101 = 0) give identical node arrays"""
    from cgal_b200 import workloads as W
    v, f = W.terrain(1024, 512)
    cases = [dict(vertices=v, faces=f), dict(triangles=W.triangle_soup(300_000)),
             dict(triangles=np.repeat(np.random.default_rng(9).uniform(0, 1, (1, 9)), 5000, 0))]
    for kw in cases:
        t = B.AABB_tree(**kw)
        t.build()
        n = t.size()
        (pn, bn), (pt, bt) = t.export_layout()[:2]
        fused = device_to_host(pn, bn).view(np.uint32).reshape(-1, 16)[:, :15].copy()
        assert hc.validate(device_to_host(pn, bn), device_to_host(pt, bt), n) == 0
        for hook, value in ((103, 1), (101, 0)):
            c = B.AABB_tree(**kw)
            c.set_option(hook, value)
            c.build()
            (qn, _), (qt, _) = c.export_layout()[:2]
            np.testing.assert_array_equal(device_to_host(qn, bn).view(np.uint32).reshape(-1, 16)[:, :15], fused)
            np.testing.assert_array_equal(device_to_host(qt, bt), device_to_host(pt, bt))
        # a refit of the fused tree (b200aabb_refit climbs with the stored links and range ends) leaves it as it is
        t.refit(**kw)
        (pn2, _) = t.export_layout()[0]
        np.testing.assert_array_equal(device_to_host(pn2, bn).view(np.uint32).reshape(-1, 16)[:, :15], fused)


def test_side_of_mesh(B, meshes):
    """Side_of_triangle_mesh (Side_of_triangle_mesh.h:265-306) through the C ABI against the oracle's restatement of the reference's ray casts: random
    points, points on the mesh (vertices, edge midpoints), points whose vertical ray runs through a vertex or an edge (re-tries), a lattice cube where
    every lattice point is a degenerate case; host and device pointers; the answer is independent of the rays (the two sides use different ones)"""
    import torch
    from test_side_of_mesh import lattice_cube, special_points
    rng = np.random.default_rng(41)
    for name in ("elephant", "finger", "cube"):
        v, f = meshes[name]
        pts = special_points(rng, v, f, 60_000 if name != "cube" else 2_000)
        g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
        want, _ = o.side_of_mesh(pts)
        got, attempts = g.side_of_mesh(pts, want_attempts=True)
        np.testing.assert_array_equal(got, want)
        assert (want == 1).sum() > 100 and (want == -1).sum() > 100 and attempts.max() >= 2 and attempts.min() == 0
        np.testing.assert_array_equal(g.side_of_mesh(torch.from_numpy(pts).cuda()).cpu().numpy(), want)
    soup = lattice_cube()
    grid = np.arange(-1, 7) * 0.5
    pts = np.array([[x, y, z] for x in grid for y in grid for z in grid])
    got = B.AABB_tree(triangles=soup).side_of_mesh(pts)
    inside = np.all((pts > 0) & (pts < 2), axis=1)
    outside = np.any((pts < 0) | (pts > 2), axis=1)
    np.testing.assert_array_equal(got, np.where(inside, 1, np.where(outside, -1, 0)))          # the cube (0,2)^3, exactly
    # unit sphere (icosphere level 5): inside / outside known analytically away from the surface
    from cgal_b200 import workloads as W
    v, f = W.icosphere(5)
    g = B.AABB_tree(v, f)
    pts = rng.uniform(-1.1, 1.1, (200_000, 3))
    r = np.linalg.norm(pts, axis=1)
    got = g.side_of_mesh(pts)
    assert np.all(got[r < 0.995] == 1) and np.all(got[r > 1.0] == -1)
    assert B.AABB_tree().side_of_mesh(pts[:5]).tolist() == [-1] * 5


def test_all_and_any_intersections_of_segments(B, meshes):
    """the intersection OBJECTS of Segment_3 queries (Intersect_3(Segment_3, Triangle_3), Segment_3_Triangle_3_intersection.h:400-548 through
    all_intersections / any_intersection, AABB_tree.h:958-985): records equal to the oracle's, objects bit for bit; segments in the plane of a flat
    grid (Segment_3 objects, clipped and collinear cases), through its vertices, ending on it; consistent with number_of_intersected_primitives"""
    from cgal_b200 import workloads as W
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    segs = W.bbox_segments(g.bbox(), 60_000)
    gv, gf = W.terrain(24, 24, amplitude=0.0)
    flat, oflat = B.AABB_tree(gv, gf), orc.OracleTree(gv, gf)
    xs = np.arange(0.0, 25.0)
    rng = np.random.default_rng(61)
    inplane = np.stack([np.full(25, -1.0), xs, np.zeros(25), np.full(25, 7.5), xs, np.zeros(25)], 1)              # along grid lines, ending inside
    oblique = np.concatenate([rng.uniform(0, 24, (200, 2)), np.zeros((200, 1)), rng.uniform(0, 24, (200, 2)), np.zeros((200, 1))], 1)   # in the plane
    through = np.stack([xs, xs[::-1], np.full(25, 5.0), xs, xs[::-1], np.full(25, -4.0)], 1)                      # through vertices
    ending = np.stack([xs, xs[::-1] * 0.5, np.full(25, 3.0), xs, xs[::-1] * 0.5, np.zeros(25)], 1)                # target on the surface
    for tree, orac, q in ((g, o, segs), (flat, oflat, np.concatenate([inplane, oblique, through, ending]))):
        off, ids, kinds, pts = tree.all_intersections(q, kind=SEG)
        roff, rids, rkinds, rpts = orac.all_intersections(q, kind=SEG)
        np.testing.assert_array_equal(off, roff)
        np.testing.assert_array_equal(np.diff(off), tree.number_of_intersected_primitives(q, SEG))
        owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
        order = np.lexsort((ids, owner)); rorder = np.lexsort((rids, owner))
        np.testing.assert_array_equal(ids[order], rids[rorder])
        np.testing.assert_array_equal(kinds[order], rkinds[rorder])
        np.testing.assert_array_equal(pts[order], rpts[rorder])
        a = tree.any_intersection(q, kind=SEG)
        cnt = np.diff(off)
        np.testing.assert_array_equal(a["prim"] == B.NO_PRIMITIVE, cnt == 0)
        lookup = {(int(w), int(i)): k for k, (w, i) in enumerate(zip(owner, ids))}
        for i in np.nonzero(cnt)[0][:4000]:
            k = lookup[(int(i), int(a["prim"][i]))]
            assert np.array_equal(a["point"][i], pts[k, :3]) and bool(a["flags"][i] & B.FLAG_SEGMENT) == (kinds[k] == 2)
    assert (kinds == 2).sum() > 100 and (kinds == 1).sum() > 50


def test_intersection_objects_of_line_and_plane_queries(B, meshes):
    """all_intersections / any_intersection for Line_3 and Plane_3 queries (Line_3_Triangle_3_intersection.h, Plane_3_Triangle_3_intersection.h) against
    the oracle: the records (query, primitive, object kind, points) are equal bit for bit; a flat grid gives in-plane lines (Segment_3 objects) and the
    plane of the grid itself (Triangle_3 objects)"""
    from cgal_b200 import workloads as W
    from test_line_plane_queries import _mesh_queries
    rng = np.random.default_rng(71)
    v, f = meshes["elephant"]
    lines, planes = _mesh_queries(rng, v, f, 4_000)
    gv, gf = W.terrain(12, 12, amplitude=0.0)
    xs = np.arange(0.0, 13.0)
    flat_lines = np.concatenate([np.stack([np.full(13, -1.0), xs, np.zeros(13), np.ones(13), np.zeros(13), np.zeros(13)], 1),          # along grid lines
                                 np.stack([xs, xs, np.full(13, 3.0), np.zeros(13), np.zeros(13), np.full(13, -1.0)], 1),                # through vertices
                                 np.concatenate([rng.uniform(0, 12, (50, 2)), np.zeros((50, 1)), rng.normal(size=(50, 2)), np.zeros((50, 1))], 1)])
    flat_planes = np.array([[0, 0, 1.0, 0], [0, 0, 1.0, -1.0], [1.0, 0, 0, -3.0], [1.0, 1.0, 0, -6.0], [1.0, 0, 0, -2.5], [0.3, 0.2, 1.0, -1.0]])
    seen = set()
    for tree_in, ql, qp in ((dict(vertices=v, faces=f), lines, planes), (dict(vertices=gv, faces=gf), flat_lines, flat_planes)):
        g, o = B.AABB_tree(**tree_in), orc.OracleTree(tree_in["vertices"], tree_in["faces"])
        for kind, q in ((B.LINE, ql), (B.PLANE, qp)):
            off, ids, kinds, pts = g.all_intersections(q, kind=kind)
            # Line_3: the reference decides nodes with the exact predicate on (point, vector) and primitives with the ROUNDED second point (a construction), so
            # where a line grazes a primitive within rounding (lines through mesh vertices) its own answer depends on its tree; this library returns the
            # tree-independent set — every primitive whose Intersect_3 is non-empty (the oracle's brute-force loop) — and the oracle's tree result is a subset
            roff, rids, rkinds, rpts = o.all_intersections(q, kind=kind, naive=(kind == B.LINE))
            if kind == B.LINE:
                toff, tids, _, _ = o.all_intersections(q, kind=kind)
                towner = np.repeat(np.arange(len(q)), np.diff(toff).astype(np.int64))
                mine = set(zip(np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64)).tolist(), ids.tolist()))
                assert set(zip(towner.tolist(), tids.tolist())) <= mine
                assert (np.diff(toff) == np.diff(off)).mean() > 0.75                            # they agree except on the grazing lines
            np.testing.assert_array_equal(off, roff)
            owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
            order = np.lexsort((ids, owner)); rorder = np.lexsort((rids, owner))
            np.testing.assert_array_equal(ids[order], rids[rorder])
            np.testing.assert_array_equal(kinds[order], rkinds[rorder])
            np.testing.assert_array_equal(pts[order], rpts[rorder])
            seen |= {(kind, int(k)) for k in np.unique(kinds)}
            a = g.any_intersection(q, kind=kind)
            cnt = np.diff(off)
            np.testing.assert_array_equal(a["prim"] == B.NO_PRIMITIVE, cnt == 0)
            lookup = {(int(w), int(i)): k for k, (w, i) in enumerate(zip(owner, ids))}
            for i in np.nonzero(cnt)[0][:2000]:
                k = lookup[(int(i), int(a["prim"][i]))]
                assert bool(a["flags"][i] & B.FLAG_SEGMENT) == (kinds[k] == 2) and bool(a["flags"][i] & B.FLAG_TRIANGLE) == (kinds[k] == 3)
                if kinds[k] in (1, 2):
                    assert np.array_equal(a["point"][i], pts[k, :3])
    assert {(B.LINE, 1), (B.LINE, 2), (B.PLANE, 1), (B.PLANE, 2), (B.PLANE, 3)} <= seen, seen


def test_sdf_shaped_ray_casting(B, meshes):
    """the ray casts of the shape-diameter function (Surface_mesh_segmentation/.../internal/SDF_calculation.h:272-356,475-515): from every facet's centroid a cone of
    25 rays about the inward normal, first_intersection with the facet itself SKIPPED (SkipPrimitiveFunctor, :40-60), as ONE batch with per-ray skip ids; the
    per-facet value (median / MAD outlier removal + weighted mean, :523-580) computed from the GPU's and from the oracle's hits is identical"""
    v, f = meshes["elephant"]
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    tri = v[f.astype(np.int64)]
    ctr = tri.mean(1)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    k = np.arange(25)                                                         # Vogel disk sampling of the cone's base, cone angle 2/3 pi
    r, th = np.sqrt((k + 0.5) / 25), k * np.pi * (3 - np.sqrt(5))
    u = np.cross(nrm, np.where(np.abs(nrm[:, :1]) < 0.9, [[1.0, 0, 0]], [[0, 1.0, 0]]))
    u /= np.linalg.norm(u, axis=1)[:, None]
    w = np.cross(nrm, u)
    half = np.pi / 3
    dirs = (-np.cos(half) * nrm[:, None, :] + np.sin(half) * (r[None, :, None] * np.cos(th)[None, :, None] * u[:, None, :] + r[None, :, None] * np.sin(th)[None, :, None] * w[:, None, :]))
    rays = np.concatenate([np.repeat(ctr[:, None, :], 25, 1), ctr[:, None, :] + dirs], 2).reshape(-1, 6)
    skip = np.repeat(np.arange(len(f), dtype=np.uint32), 25)
    got = g.first_intersection_skip(rays, skip_ids=skip)
    ref = o.first_intersection_skip(rays, skip_ids=skip)
    flagged = (got["flags"] & (B.FLAG_TIE | B.FLAG_NEAR_TIE)) != 0
    same = got["prim"] == ref["prim"]
    assert np.all(same | flagged) and np.all(got["prim"] != skip)
    np.testing.assert_array_equal(got["t"][same], ref["t"][same])
    assert (got["prim"] != B.NO_PRIMITIVE).mean() > 0.9                        # a closed mesh: inward rays hit

    def sdf(t, prim):
        d = np.where(prim != B.NO_PRIMITIVE, t * np.linalg.norm(dirs.reshape(-1, 3), axis=1), np.nan).reshape(-1, 25)
        med = np.nanmedian(d, axis=1)
        mad = np.nanmedian(np.abs(d - med[:, None]), axis=1)
        keep = np.abs(d - med[:, None]) <= 1.5 * mad[:, None]
        wgt = np.where(keep & ~np.isnan(d), 1.0, 0.0)
        return np.nansum(np.where(wgt > 0, d, 0.0) * wgt, axis=1) / np.maximum(wgt.sum(1), 1)
    np.testing.assert_array_equal(sdf(got["t"], got["prim"])[same.reshape(-1, 25).all(1)], sdf(ref["t"], ref["prim"])[same.reshape(-1, 25).all(1)])


def test_ray_kernels_agree(B, meshes):
    """first_intersection with persistent threads and ray replacement (B200AABB_OPT_RAY_KERNEL = 1) returns the records of the thread-per-ray kernel, bit for bit,
    flags and traversal counters included"""
    import torch
    from cgal_b200 import workloads as W
    from cgal_b200.aabb_tree import OPT_RAY_KERNEL
    v, f = meshes["elephant"]
    g = B.AABB_tree(v, f)
    gv, gf = W.terrain(24, 24, amplitude=0.0)
    flat = B.AABB_tree(gv, gf)
    xs = np.arange(0.0, 25.0)
    grid_rays = np.concatenate([np.stack([np.full(25, -1.0), xs, np.zeros(25), np.zeros(25), xs, np.zeros(25)], 1), np.stack([xs, xs[::-1], np.full(25, 5.0), xs, xs[::-1], np.full(25, 4.0)], 1)])
    for tree, rays in ((g, W.sphere_rays(g.bbox(), 100_000)), (g, W.sphere_rays(g.bbox(), 37)), (flat, grid_rays)):
        tree.set_option(OPT_RAY_KERNEL, 0)
        a = tree.first_intersection(rays)
        sa = tree.stats()
        tree.set_option(OPT_RAY_KERNEL, 1)
        b = tree.first_intersection(rays)
        sb = tree.stats()
        d = tree.first_intersection(torch.from_numpy(rays).cuda())
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
            np.testing.assert_array_equal(a[k], d[k].cpu().numpy().view(a[k].dtype))          # (torch has no uint32: the device ids come back as int32)
        assert (sa["nodes_visited"], sa["leaves_tested"], sa["ties"]) == (sb["nodes_visited"], sb["leaves_tested"], sb["ties"])
        tree.set_option(OPT_RAY_KERNEL, 0)
    # the same scheduling for do_intersect / counts / any_intersected_primitive (intersect_stream_kernel): Segment_3, Ray_3 and Line_3 queries
    segs = W.bbox_segments(g.bbox(), 50_000)
    lines = np.concatenate([segs[:, :3], segs[:, 3:] - segs[:, :3]], 1)
    for kind, q in ((B.SEGMENT, segs), (B.RAY, segs), (B.LINE, lines)):
        g.set_option(OPT_RAY_KERNEL, 0)
        want = (g.do_intersect(q, kind), g.number_of_intersected_primitives(q, kind))
        g.set_option(OPT_RAY_KERNEL, 1)
        np.testing.assert_array_equal(g.do_intersect(q, kind), want[0])
        np.testing.assert_array_equal(g.number_of_intersected_primitives(q, kind), want[1])
        np.testing.assert_array_equal(g.number_of_intersected_primitives(torch.from_numpy(q).cuda(), kind).cpu().numpy().view(np.uint32), want[1])
        a = g.any_intersected_primitive(q, kind)
        np.testing.assert_array_equal(a == B.NO_PRIMITIVE, want[1] == 0)
        g.set_option(OPT_RAY_KERNEL, 0)


@pytest.mark.parametrize("name", ["elephant", "cube"])
def test_triangle_queries(B, meshes, name):
    """Triangle_3 queries (kind 4: three vertices per query) — do_intersect, number_of_intersected_primitives, any_ / all_intersected_primitives — against the
    oracle's restatement of Triangle_3_Triangle_3_do_intersect.h (pinned by the reference's KATs, tests/test_triangle_queries.py): random triangles, copies of
    facets (coplanar with their primitive), facets slid inside their plane, triangles through mesh vertices; host and device pointers, sorted and unsorted"""
    import torch
    from test_triangle_queries import mesh_triangles
    v, f = meshes[name]
    rng = np.random.default_rng(44)
    q = mesh_triangles(rng, v, f, 20_000 if name == "elephant" else 2_000)
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    want = o.count(q, B.TRIANGLE)
    assert want.max() > 2 and (want == 0).any()
    for sort in (1, 0):
        g.set_option(B.OPT_SORT_QUERIES, sort)
        np.testing.assert_array_equal(g.number_of_intersected_primitives(q, B.TRIANGLE), want)
        np.testing.assert_array_equal(g.do_intersect(q, B.TRIANGLE) != 0, want > 0)
    off, ids = g.all_intersected_primitives(q, B.TRIANGLE)
    roff, rids = o.all_intersected_primitives(q, B.TRIANGLE)
    np.testing.assert_array_equal(off, roff)
    owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
    key = lambda own, i: np.sort(own.astype(np.uint64) << np.uint64(32) | i.astype(np.uint64))   # noqa: E731
    np.testing.assert_array_equal(key(owner, ids), key(owner, rids))
    a = g.any_intersected_primitive(q, B.TRIANGLE)
    np.testing.assert_array_equal(a == B.NO_PRIMITIVE, want == 0)
    pairs = set(zip(owner.tolist(), ids.tolist()))
    assert all((i, int(a[i])) in pairs for i in np.nonzero(want)[0][:3000])
    qd = torch.from_numpy(q).cuda()
    np.testing.assert_array_equal(g.number_of_intersected_primitives(qd, B.TRIANGLE).cpu().numpy(), want)
    np.testing.assert_array_equal(g.do_intersect(qd, B.TRIANGLE).cpu().numpy() != 0, want > 0)
    with pytest.raises(B.B200AABBError):                                    # the intersection OBJECTS take kinds 0..3 only
        g.all_intersections(q[:4], B.TRIANGLE)


def test_tree_against_tree(B, meshes):
    """tree.do_intersect(other_tree) (AABB_traits_3.h:679-690) against the oracle's traversal with the other tree as the query: overlapping, touching and
    separated copies of a mesh, trees of different sizes in both argument orders, an empty tree, and the witness pair checked with the exact predicate"""
    v, f = meshes["elephant"]
    cv, cf = meshes["cube"]
    ext = (v.max(0) - v.min(0))
    g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
    soup = soup_of(v, f)
    n_true = 0
    for shift in ([0.0, 0.0, 0.0], [0.3 * ext[0], 0.0, 0.0], [0.9 * ext[0], 0.1 * ext[1], 0.0], [1.001 * ext[0], 0.0, 0.0], [0.0, 0.0, 5.0 * ext[2]]):
        w = v + np.array(shift)
        g2, o2 = B.AABB_tree(w, f), orc.OracleTree(w, f)
        want = o.do_intersect_tree(o2)
        assert o2.do_intersect_tree(o) == want
        n_true += want
        hit, ia, ib = g.do_intersect_tree(g2, witness=True)
        assert hit == want, shift
        assert g2.do_intersect_tree(g) == want, shift
        if hit:
            assert orc.tri_tri_do_intersect(soup[ia], soup_of(w, f)[ib]) == 1
        else:
            assert ia == B.NO_PRIMITIVE and ib == B.NO_PRIMITIVE
    assert 2 <= n_true <= 4
    # a small tree against a large one: a cube placed inside the bounding box of the elephant, then far away
    ctr = (v.max(0) + v.min(0)) / 2
    cube = (cv - cv.mean(0)) / (cv.max() - cv.min()) * 0.5 * ext.min()
    for place, far in ((ctr, False), (ctr + 10 * ext, True)):
        c, oc = B.AABB_tree(cube + place, cf), orc.OracleTree(cube + place, cf)
        want = o.do_intersect_tree(oc)
        assert g.do_intersect_tree(c) == want and c.do_intersect_tree(g) == want
        if far:
            assert not want
    e = B.AABB_tree()
    assert g.do_intersect_tree(e) is False and e.do_intersect_tree(g) is False


def test_tree_against_tree_under_transformations(B, meshes):
    """Rigid_triangle_mesh_collision_detection::does_A_intersect_B: both trees built once, each mesh under a rotation + scaling + translation given per call;
    against the oracle's traversal with transformed datums and boxes (AABB_traversal_traits_with_transformation.h) and, on the cube, against brute force over
    all pairs of transformed primitives; the witness pair is re-checked with the exact predicate on the transformed triangles"""
    from test_triangle_queries import rigid
    rng = np.random.default_rng(47)
    for name, trials in (("cube", 40), ("elephant", 24)):
        v, f = meshes[name]
        ext = (v.max(0) - v.min(0)).max()
        g, g2, o, o2 = B.AABB_tree(v, f), B.AABB_tree(v, f), orc.OracleTree(v, f), orc.OracleTree(v, f)
        soup = soup_of(v, f)
        n_true = 0
        for it in range(trials):
            xa = rigid(rng, shift=rng.uniform(-1, 1, 3) * ext) if it % 2 else None
            xb = rigid(rng, scale=rng.uniform(0.5, 1.5), shift=rng.uniform(-1.2, 1.2, 3) * ext * (0.3 if it % 3 == 0 else 1.0))
            want = o.do_intersect_tree_transformed(xa, o2, xb, naive=(name == "cube"))
            n_true += want
            hit, ia, ib = g.do_intersect_tree(g2, witness=True, xf=xa, other_xf=xb)
            assert hit == want, (name, it)
            assert g2.do_intersect_tree(g, xf=xb, other_xf=xa) == want, (name, it)
            if hit:
                ta = soup[ia] if xa is None else hc.xf_apply(xa, soup[ia])
                assert orc.tri_tri_do_intersect(ta, hc.xf_apply(xb, soup[ib])) == 1
        assert 3 < n_true < trials - 3, (name, n_true)
        # a tree against ITSELF under another transformation (one handle for both arguments)
        far = np.hstack([np.eye(3), np.array([[10.0 * ext], [0.0], [0.0]])])
        assert g.do_intersect_tree(g) and not g.do_intersect_tree(g, other_xf=far) and not g.do_intersect_tree(g, xf=far) and g.do_intersect_tree(g, xf=far, other_xf=far)


def test_side_of_mesh_under_a_transformation(B, meshes):
    """b200aabb_side_of_mesh_transformed (does_A_contains_a_CC_of_B of Rigid_triangle_mesh_collision_detection.h): the tree built once on the mesh, the affine
    map given per call, against the oracle on a tree built over the rounded images of the primitives; host and device pointers"""
    import torch
    from test_side_of_mesh import transformed_soup
    from test_triangle_queries import rigid
    rng = np.random.default_rng(48)
    for name in ("cube", "elephant"):
        v, f = meshes[name]
        soup = soup_of(v, f)
        g = B.AABB_tree(v, f)
        for trial in range(3):
            m = rigid(rng, scale=rng.uniform(0.5, 2.0), shift=rng.uniform(-3, 3, 3)) if trial else np.array([[0, -1, 0, 0.5], [1, 0, 0, -0.25], [0, 0, 1, 2.0]], float)
            img = transformed_soup(m, soup)
            w = img.reshape(-1, 3)
            lo, hi = w.min(0), w.max(0)
            pts = np.concatenate([lo + rng.uniform(-0.1, 1.1, (20_000, 3)) * (hi - lo), w[rng.integers(0, len(w), 500)],
                                  (w[rng.integers(0, len(w), 500)] + w[rng.integers(0, len(w), 500)]) / 2])
            want, _ = orc.OracleTree(soup=img).side_of_mesh(pts)
            np.testing.assert_array_equal(g.side_of_mesh(pts, xf=m), want)
            np.testing.assert_array_equal(g.side_of_mesh(torch.from_numpy(pts).cuda(), xf=m).cpu().numpy(), want)
            assert (want == 1).any() and (want == -1).any() and (want == 0).any()
        ident = np.hstack([np.eye(3), np.zeros((3, 1))])
        np.testing.assert_array_equal(g.side_of_mesh(pts, xf=ident), g.side_of_mesh(pts))


def test_segment_primitives(B, meshes):
    """a tree over Segment_3 primitives (the edges of the elephant, b200aabb_set_segments): closest_point / squared_distance bit-equal to the oracle's traversal with
    the reference's Construct_projected_point_3(Segment_3, Point_3), Plane_3 and Triangle_3 id queries, max_distance_to_mesh; other entry points are refused"""
    import torch
    from test_segment_primitives import mesh_edges
    rng = np.random.default_rng(53)
    v, f = meshes["elephant"]
    segs = mesh_edges(v, f)
    g, o = B.AABB_tree(segments=segs), orc.OracleTree(segments=segs)
    assert g.size() == len(segs)
    np.testing.assert_array_equal(g.bbox(), o.bbox())
    lo, hi = v.min(0), v.max(0)
    pts = np.concatenate([lo + rng.uniform(-0.2, 1.2, (30_000, 3)) * (hi - lo), segs[rng.integers(0, len(segs), 500), :3],
                          (segs[rng.integers(0, len(segs), 500), :3] + segs[rng.integers(0, len(segs), 500), 3:]) / 2])
    got, want = g.closest_point_and_primitive(pts), o.closest_point(pts)
    np.testing.assert_array_equal(got["sqdist"], want["sqd"])
    same = got["prim"] == want["prim"]
    np.testing.assert_array_equal(got["point"][same], want["cp"][same])
    # (on the edges of a mesh the closest point is a vertex for about half of the queries: an exact tie between the edges that share it, which is reported)
    assert ((got["flags"][~same] & B.FLAG_TIE) != 0).all() and same.mean() > 0.3
    gd = g.closest_point_and_primitive(torch.from_numpy(pts).cuda())
    np.testing.assert_array_equal(gd["sqdist"].cpu().numpy(), want["sqd"])
    md, far = g.max_distance_to_mesh(pts)
    assert md == np.sqrt(want["sqd"].max()) and want["sqd"][far] == want["sqd"].max()
    nrm = rng.normal(size=(5000, 3)); x = lo + rng.uniform(0, 1, (5000, 3)) * (hi - lo)
    planes = np.concatenate([nrm, -np.sum(nrm * x, axis=1, keepdims=True)], axis=1)
    planes[::4, :3] = np.eye(3)[rng.integers(0, 3, 1250)]; planes[::4, 3] = -segs[rng.integers(0, len(segs), 1250), 0]
    ctr = lo + rng.uniform(0, 1, (5000, 3)) * (hi - lo)
    tris = np.concatenate([ctr, ctr + rng.uniform(-0.2, 0.2, (5000, 3)) * (hi - lo), ctr + rng.uniform(-0.2, 0.2, (5000, 3)) * (hi - lo)], axis=1)
    tris[::5] = v[f[rng.integers(0, len(f), 1000)]].reshape(-1, 9)
    for kind, q in ((B.PLANE, planes), (B.TRIANGLE, tris)):
        wantc = o.count(q, kind)
        np.testing.assert_array_equal(g.number_of_intersected_primitives(q, kind), wantc)
        np.testing.assert_array_equal(g.do_intersect(q, kind) != 0, wantc > 0)
        off, ids = g.all_intersected_primitives(q, kind)
        roff, rids = o.all_intersected_primitives(q, kind)
        np.testing.assert_array_equal(off, roff)
        owner = np.repeat(np.arange(len(q)), np.diff(off).astype(np.int64))
        key = lambda own, i: np.sort(own.astype(np.uint64) << np.uint64(32) | i.astype(np.uint64))   # noqa: E731
        np.testing.assert_array_equal(key(owner, ids), key(owner, rids))
        assert wantc.max() > 2 and (wantc == 0).any()
    for refused in (lambda: g.do_intersect(segs[:4], B.SEGMENT), lambda: g.first_intersection(segs[:4]), lambda: g.side_of_mesh(pts[:4]),
                    lambda: g.do_intersect_tree(g), lambda: g.all_intersections(planes[:4], B.PLANE)):
        with pytest.raises(B.B200AABBError):
            refused()
    g.insert(vertices=v, faces=f)                                           # back to triangles: the flag goes with the primitives
    assert g.do_intersect(segs[:4], B.SEGMENT).shape == (4,)


def _off_text(v, f, fmt="%.17g", header="OFF\n%d %d 0\n", sep=" ", eol="\n"):
    lines = [header % (len(v), len(f))]
    lines += [sep.join(fmt % c for c in p) + eol for p in v]
    lines += ["3" + sep + sep.join(str(int(i)) for i in t) + eol for t in f]
    return "".join(lines).encode()


def test_off_ingest_on_the_device(B, meshes):
    """b200aabb_set_mesh_off: the text of an OFF file tokenised and parsed by kernels; the tree must be the one built from the same file read on the host
    (coordinates = Python's correctly rounded float() of the literals, facets in file order): equal boxes and bit-equal query results; malformed and
    unsupported files are refused with the tree left empty"""
    rng = np.random.default_rng(71)
    v, f = meshes["elephant"]
    variants = [dict(), dict(fmt="%.9g"), dict(fmt="%.6f", sep="  ", eol="\r\n"), dict(fmt="%.20e"), dict(header="# a comment\nOFF # keyword\n# counts follow\n%d %d 0\n", sep="\t")]
    bb = None
    for kw in variants:
        text = _off_text(v, f, **kw)
        vv = np.array([[float(kw.get("fmt", "%.17g") % c) for c in p] for p in v])          # what a correctly rounding reader gets
        g, h = B.AABB_tree(), B.AABB_tree(vv, f)
        g.insert_off(text)
        assert g.size() == len(f)
        np.testing.assert_array_equal(g.bbox(), h.bbox())
        bb = h.bbox()
        from cgal_b200 import workloads as W
        rays, pts = W.sphere_rays(bb, 20_000), W.bbox_points(bb, 20_000)
        a, b = g.first_intersection(rays), h.first_intersection(rays)
        for k in ("prim", "t", "point", "flags"):
            np.testing.assert_array_equal(a[k], b[k])
        a, b = g.closest_point_and_primitive(pts), h.closest_point_and_primitive(pts)
        for k in ("prim", "sqdist", "point", "flags"):
            np.testing.assert_array_equal(a[k], b[k])
    # 25-digit literals: the ones whose dropped digits decide are parsed by strtod on the host text; the result is still the correctly rounded value
    long_v = np.array([[float("%.25e" % c) for c in p] for p in v])
    g = B.AABB_tree()
    g.insert_off(_off_text(v, f, fmt="%.25e"))
    np.testing.assert_array_equal(g.bbox(), B.AABB_tree(long_v, f).bbox())
    o = orc.OracleTree(long_v, f)
    pts = W.bbox_points(bb, 3000)
    check_closest_point(g.closest_point_and_primitive(pts), o.closest_point(pts), o, pts)
    # refused: quads, colours, comments in the body, a bad number, a vertex index out of range, a wrong keyword, missing tokens
    good = _off_text(v[:4], np.array([[0, 1, 2], [0, 2, 3]], np.uint32))
    bad = [b"OFF\n4 1 0\n0 0 0\n1 0 0\n1 1 0\n0 1 0\n4 0 1 2 3\n", good.replace(b"3 0 1 2\n", b"3 0 1 2 255 0 0\n"), good.replace(b"3 0 2 3", b"# c\n3 0 2 3"),
           good.replace(b"3 0 1 2", b"3 0 1 x"), good.replace(b"3 0 2 3", b"3 0 2 4"), good.replace(b"OFF", b"COFF"), good[:-8], b"OFF\n", b""]
    for text in bad:
        t = B.AABB_tree(v, f)
        with pytest.raises(B.B200AABBError):
            t.insert_off(text)
        assert t.size() == 0                                               # a failed ingest leaves an empty tree
    t = B.AABB_tree()
    t.insert_off(b"OFF\n3 0 0\n0 0 0\n1 0 0\n0 1 0\n")                     # no facets: an empty tree
    assert t.size() == 0
