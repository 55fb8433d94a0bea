"""-m gpu: the DEVICE-compiled predicates and constructions (static filter + on-device exact stage) against the oracle,
record by record, on random, lattice-degenerate, nearly-coplanar and ulp-nudged inputs.

tests/test_device_logic_on_host.py checks the same headers compiled for the host; this file exists because the device
compilation is a different compiler: cicc once mis-merged the stack slots of the force-inlined expansion routines and the
GPU returned wrong exact signs for inputs the host build got right (see cgal_b200/csrc/exact.cuh)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dbg():
    from cgal_b200.aabb_tree import debug_predicate
    return debug_predicate


def ulp_nudge(rng, x, k=2):
    x = np.array(x, dtype=np.float64)
    for _ in range(k):
        step = rng.integers(-1, 2, size=x.shape)
        x = np.where(step > 0, np.nextafter(x, np.inf), np.where(step < 0, np.nextafter(x, -np.inf), x))
    return x


def point_sets(rng, n, npts):
    """(n, npts, 3) points; the LAST point is tied to the first three in several degenerate ways"""
    out = np.empty((n, npts, 3))
    for i in range(n):
        mode = i % 6
        scale = [1.0, 10.0, 4096.0][i % 3]
        if mode == 0:
            p = rng.uniform(-scale, scale, size=(npts, 3))
        elif mode == 1:    # integer lattice: many exactly degenerate configurations
            p = rng.integers(-3, 4, size=(npts, 3)).astype(float)
        elif mode == 2:    # last point = affine combination of the first three: coplanar up to rounding
            p = rng.uniform(-scale, scale, size=(npts, 3))
            w = rng.uniform(-1, 2, size=3)
            w /= w.sum()
            p[-1] = w @ p[:3]
        elif mode == 3:    # lattice + large offset, nudged by a few ulps: catastrophic cancellation
            p = rng.integers(-3, 4, size=(npts, 3)).astype(float) + 1024.0
            p[-1] = ulp_nudge(rng, p[-1])
        elif mode == 4:    # terrain-like: unit grid in xy, small z, far query point
            p = np.concatenate([rng.integers(0, 4096, size=(npts, 2)).astype(float), rng.uniform(-40, 40, size=(npts, 1))], 1)
            p[-1] = p[0] + rng.uniform(0, 1) * (p[1] - p[0]) + rng.uniform(0, 1) * (p[2] - p[0])
        else:              # last point on the segment of the first two
            p = rng.uniform(-scale, scale, size=(npts, 3))
            p[-1] = p[0] + rng.uniform(-0.5, 1.5) * (p[1] - p[0])
        out[i] = p
    return out


def test_device_orientation_filter_and_exact_stage(dbg):
    rng = np.random.default_rng(7)
    pts = point_sets(rng, 20000, 4)
    want = np.array([orc.orientation(*p, exact=True) for p in pts])
    filt, used, _ = dbg(0, pts.reshape(len(pts), 12))
    exact, _, _ = dbg(8, pts.reshape(len(pts), 12))
    np.testing.assert_array_equal(exact, want)      # the exact stage alone decides every input
    np.testing.assert_array_equal(filt, want)       # filter + exact fallback
    assert (used > 0).sum() > 2000 and (want == 0).sum() > 100


def test_device_ray_direction_and_coplanar_orientation(dbg):
    rng = np.random.default_rng(8)
    pts = point_sets(rng, 6000, 5)
    want = np.array([orc.orientation_dir(*p) for p in pts])
    got, _, _ = dbg(1, pts.reshape(len(pts), 15))
    np.testing.assert_array_equal(got, want)
    pts = point_sets(rng, 6000, 3)
    pts[::2, :, 2] = 0.0                            # genuinely planar configurations
    want = np.array([orc.coplanar_orientation(*p) for p in pts])
    got, _, _ = dbg(2, pts.reshape(len(pts), 9))
    np.testing.assert_array_equal(got, want)
    assert (want == 0).sum() > 5


def test_device_compare_distance(dbg):
    rng = np.random.default_rng(9)
    n = 12000
    p = rng.uniform(-50, 50, size=(n, 3))
    q = rng.uniform(-50, 50, size=(n, 3))
    r = rng.uniform(-50, 50, size=(n, 3))
    k = n // 3
    r[:k] = 2 * p[:k] - q[:k]                       # mirror image: equal distance up to rounding
    r[k:2 * k] = ulp_nudge(rng, q[k:2 * k])         # the same point nudged by ulps
    r[2 * k:2 * k + 500] = q[2 * k:2 * k + 500]     # identical points
    rec = np.concatenate([p, q, r], 1)
    want = np.array([orc.compare_distance(a[:3], a[3:6], a[6:]) for a in rec])
    got, used, _ = dbg(3, rec)
    np.testing.assert_array_equal(got, want)
    assert (used > 0).sum() > 1000 and (want == 0).sum() >= 500


def tri_line_records(rng, n):
    rec = np.empty((n, 15))
    for i in range(n):
        mode = i % 5
        scale = [1.0, 100.0, 4096.0][i % 3]
        tri = rng.uniform(-scale, scale, size=(3, 3)) if mode != 1 else rng.integers(-3, 4, size=(3, 3)).astype(float)
        if mode == 0:
            s, q = rng.uniform(-scale, scale, size=(2, 3))
        elif mode == 1:    # lattice: coplanar, through vertices, along edges
            s, q = rng.integers(-3, 4, size=(2, 3)).astype(float)
            if np.all(s == q):
                q = q + 1.0
        elif mode == 2:    # through a point of an edge (up to rounding)
            e = tri[0] + rng.uniform(0, 1) * (tri[1] - tri[0])
            s = rng.uniform(-scale, scale, size=3)
            q = s + rng.uniform(0.3, 2.0) * (e - s)
        elif mode == 3:    # through a vertex exactly (second point IS the vertex)
            s = rng.uniform(-scale, scale, size=3)
            q = tri[int(rng.integers(0, 3))].copy()
        else:              # in the supporting plane (up to rounding)
            w1, w2 = rng.uniform(-1, 2, size=(2, 3))
            s, q = (w1 / w1.sum()) @ tri, (w2 / w2.sum()) @ tri
        rec[i, :9], rec[i, 9:12], rec[i, 12:] = tri.ravel(), s, q
    return rec


def test_device_triangle_tests_and_hit_construction(dbg):
    rng = np.random.default_rng(10)
    rec = tri_line_records(rng, 12000)
    for which, is_ray in ((4, 0), (5, 1)):
        want = np.array([orc.tri_do_intersect(r[:9], r[9:], is_ray) for r in rec])
        got, _, _ = dbg(which, rec)
        np.testing.assert_array_equal(got, want)
        assert 0.05 < want.mean() < 0.95
    kinds, _, val = dbg(6, rec)
    for i, r in enumerate(rec):
        k, pts = orc.tri_ray_intersection(r[:9], r[9:])
        assert k == kinds[i]
        if k >= 1:
            np.testing.assert_array_equal(val[i, 1:4], pts[0])      # bit-exact Point_3 / first endpoint
        if k == 2:
            np.testing.assert_array_equal(val[i, 4:7], pts[1])
    assert (kinds == 2).sum() > 10 and (kinds == 1).sum() > 1000


def test_device_projection_is_bit_exact(dbg):
    rng = np.random.default_rng(11)
    n = 20000
    tri = rng.uniform(-10, 10, size=(n, 9))
    tri[::4] = rng.integers(-3, 4, size=(len(tri[::4]), 9)).astype(float)
    q = rng.uniform(-30, 30, size=(n, 3))
    rec = np.concatenate([tri, q], 1)
    ok = np.array([np.any(np.cross(t[3:6] - t[:3], t[6:] - t[:3])) for t in tri])      # skip exactly degenerate lattice triangles
    rec = rec[ok]
    _, _, val = dbg(7, rec)
    want = np.array([orc.project_on_triangle(r[:9], r[9:]) for r in rec])
    np.testing.assert_array_equal(val[:, :3], want)
