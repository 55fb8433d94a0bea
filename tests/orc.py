"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [_dp, C.c_uint64, _u32p, C.c_uint64]
        L.orc_create_soup.restype = C.c_void_p
        L.orc_create_soup.argtypes = [_dp, C.c_uint64]
        L.orc_create_segments.restype = C.c_void_p
        L.orc_create_segments.argtypes = [_dp, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_build_seconds.restype = C.c_double
        L.orc_build_seconds.argtypes = [C.c_void_p]
        L.orc_kd_seconds.restype = C.c_double
        L.orc_kd_seconds.argtypes = [C.c_void_p]
        L.orc_size.restype = C.c_uint64
        L.orc_size.argtypes = [C.c_void_p]
        L.orc_accelerate_distance_queries.argtypes = [C.c_void_p]
        L.orc_do_not_accelerate_distance_queries.argtypes = [C.c_void_p]
        L.orc_bbox.argtypes = [C.c_void_p, _dp]
        L.orc_do_intersect.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, _u8p, C.c_int]
        L.orc_count.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, _u32p, C.c_int]
        L.orc_any_intersected_primitive.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, _u32p, C.c_int]
        L.orc_all_intersected_primitives.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), _u32p, C.c_int]
        L.orc_all_intersected_primitives.restype = C.c_int
        L.orc_ray_intersections.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, _u32p, C.POINTER(C.c_uint64), _u32p, _u8p, _dp, C.c_int]
        L.orc_ray_intersections.restype = C.c_int
        L.orc_first_intersection.argtypes = [C.c_void_p, _dp, C.c_uint64, _u32p, _dp, _dp, _u8p, _u8p, C.c_int]
        L.orc_first_intersection_skip.argtypes = [C.c_void_p, _dp, C.c_uint64, _u32p, _u32p, C.c_uint32, _u32p, _dp, C.c_int]
        L.orc_closest_point.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp, _u32p, _dp, C.c_int]
        L.orc_closest_point.restype = C.c_int
        L.orc_naive_count.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, _u32p, C.c_int]
        L.orc_naive_closest.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp, _u32p, _u32p, C.c_int]
        L.orc_naive_first_intersection.argtypes = [C.c_void_p, _dp, C.c_uint64, _u32p, _dp, _u32p, C.c_int]
        L.orc_candidate_closest.argtypes = [C.c_void_p, C.c_uint32, _dp, _dp]
        L.orc_candidate_ray_hit.argtypes = [C.c_void_p, C.c_uint32, _dp, _dp, _dp]
        L.orc_candidate_ray_hit.restype = C.c_int
        for name in ("orc_orientation", "orc_orientation_exact", "orc_orientation_dir", "orc_coplanar_orientation3",
                     "orc_coplanar_orientation4", "orc_compare_distance", "orc_compare_distance_exact"):
            getattr(L, name).argtypes = [_dp]
            getattr(L, name).restype = C.c_int
        L.orc_bbox_do_intersect.argtypes = [_dp, _dp, C.c_int]
        L.orc_bbox_do_intersect_exact.argtypes = [_dp, _dp, C.c_int]
        L.orc_tri_do_intersect.argtypes = [_dp, _dp, C.c_int]
        L.orc_tri_tri_do_intersect.argtypes = [_dp, _dp]
        L.orc_do_intersect_tree.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_do_intersect_tree_transformed.argtypes = [C.c_void_p, _dp, C.c_void_p, _dp]
        L.orc_naive_do_intersect_transformed.argtypes = [C.c_void_p, _dp, C.c_void_p, _dp]
        L.orc_tri_ray_intersection.argtypes = [_dp, _dp, _dp]
        L.orc_project_on_triangle.argtypes = [_dp, _dp, _dp]
        L.orc_ball_reaches_box.argtypes = [_dp, _dp, _dp]
        L.orc_intersection_distance.argtypes = [_dp, _dp, _dp]
        L.orc_stats.argtypes = [C.POINTER(C.c_uint64)]
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def traversal_totals():
    """(child boxes tested, primitives tested) by the batched oracle queries since the last call"""
    out = np.zeros(2, np.uint64)
    lib().orc_traversal_totals(out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return int(out[0]), int(out[1])


def nthreads():
    return max(1, os.cpu_count() or 1)


_QDOUBLES = {3: 4, 4: 9}   # doubles per query: Plane_3 (a, b, c, d), Triangle_3 (three vertices); every other kind 6


class OracleTree:
    """Mirror of CGAL::AABB_tree<AABB_traits_3<Epick, AABB_face_graph_triangle_primitive<Surface_mesh>>> (CPU)."""

    def __init__(self, verts=None, faces=None, soup=None, accelerate=True, segments=None):
        L = lib()
        if segments is not None:                                            # Segment_3 primitives: (n, 6) = (source, target)
            self.segments = _d(segments).reshape(-1, 6)
            self.h = L.orc_create_segments(_ptr(self.segments, _dp), len(self.segments))
        elif soup is not None:
            self.soup = _d(soup).reshape(-1, 9)
            self.h = L.orc_create_soup(_ptr(self.soup, _dp), len(self.soup))
        else:
            self.verts = _d(verts).reshape(-1, 3)
            self.faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
            self.h = L.orc_create(_ptr(self.verts, _dp), len(self.verts), _ptr(self.faces, _u32p), len(self.faces))
        if accelerate and self.size() > 0:
            L.orc_accelerate_distance_queries(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def size(self):
        return lib().orc_size(self.h)

    def build_seconds(self):
        return lib().orc_build_seconds(self.h)

    def kd_seconds(self):
        return lib().orc_kd_seconds(self.h)

    def bbox(self):
        out = np.zeros(6)
        lib().orc_bbox(self.h, _ptr(out, _dp))
        return out

    def do_intersect(self, q, kind, threads=None):
        q = _d(q).reshape(-1, _QDOUBLES.get(kind, 6))
        out = np.zeros(len(q), np.uint8)
        lib().orc_do_intersect(self.h, _ptr(q, _dp), len(q), kind, _ptr(out, _u8p), threads or nthreads())
        return out

    def count(self, q, kind, threads=None, naive=False):
        q = _d(q).reshape(-1, _QDOUBLES.get(kind, 6))
        out = np.zeros(len(q), np.uint32)
        f = lib().orc_naive_count if naive else lib().orc_count
        f(self.h, _ptr(q, _dp), len(q), kind, _ptr(out, _u32p), threads or nthreads())
        return out

    def do_intersect_tree(self, other):
        """tree.do_intersect(other_tree), AABB_traits_3.h:679-690"""
        return bool(lib().orc_do_intersect_tree(self.h, other.h))

    def do_intersect_tree_transformed(self, xf, other, other_xf, naive=False):
        """Rigid_triangle_mesh_collision_detection::does_A_intersect_B: both trees under 3 x 4 affine maps (12 doubles, row major; None = identity)"""
        ident = np.array([1.0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0])
        a = _d(ident if xf is None else xf).ravel()
        b = _d(ident if other_xf is None else other_xf).ravel()
        f = lib().orc_naive_do_intersect_transformed if naive else lib().orc_do_intersect_tree_transformed
        return bool(f(self.h, _ptr(a, _dp), other.h, _ptr(b, _dp)))

    def side_of_mesh(self, pts, threads=None):
        """Side_of_triangle_mesh: (side[n] in {-1 outside, 0 boundary, 1 inside}, rays cast per point)"""
        pts = _d(pts).reshape(-1, 3)
        side, att = np.zeros(len(pts), np.int8), np.zeros(len(pts), np.uint8)
        lib().orc_side_of_mesh(self.h, _ptr(pts, _dp), len(pts), side.ctypes.data_as(C.POINTER(C.c_int8)), _ptr(att, _u8p), threads or nthreads())
        return side, att

    def any_intersected_primitive(self, q, kind, threads=None):
        q = _d(q).reshape(-1, _QDOUBLES.get(kind, 6))
        out = np.zeros(len(q), np.uint32)
        lib().orc_any_intersected_primitive(self.h, _ptr(q, _dp), len(q), kind, _ptr(out, _u32p), threads or nthreads())
        return out

    def all_intersected_primitives(self, q, kind, threads=None):
        """(offsets[n+1], ids): ids of query i, in the reference's traversal order, are ids[offsets[i]:offsets[i+1]]"""
        q = _d(q).reshape(-1, _QDOUBLES.get(kind, 6))
        offsets = np.zeros(len(q) + 1, np.uint64)
        np.cumsum(self.count(q, kind, threads), out=offsets[1:])
        ids = np.zeros(max(int(offsets[-1]), 1), np.uint32)
        rc = lib().orc_all_intersected_primitives(self.h, _ptr(q, _dp), len(q), kind, offsets.ctypes.data_as(C.POINTER(C.c_uint64)), _ptr(ids, _u32p), threads or nthreads())
        assert rc == 0
        return offsets, ids[:int(offsets[-1])]

    def all_intersections(self, rays, first_only=False, threads=None, kind=1, naive=False):
        """all_intersections(Ray_3 | kind 0 Segment_3 | 2 Line_3 | 3 Plane_3) (first_only: any_intersection; naive: every primitive tested, no tree):
        (offsets[n+1], ids, kinds (1 Point_3 / 2 Segment_3 / 3 the primitive's Triangle_3), pts[total,6])"""
        if naive:
            first_only = 2
        q = _d(rays).reshape(-1, _QDOUBLES.get(kind, 6))
        th = threads or nthreads()
        counts = np.zeros(len(q), np.uint32)
        f = lib().orc_line_intersections
        f(self.h, _ptr(q, _dp), len(q), int(kind), int(first_only), _ptr(counts, _u32p), None, None, None, None, th)
        offsets = np.zeros(len(q) + 1, np.uint64)
        np.cumsum(counts, out=offsets[1:])
        tot = max(int(offsets[-1]), 1)
        ids, kinds, pts = np.zeros(tot, np.uint32), np.zeros(tot, np.uint8), np.zeros((tot, 6))
        rc = f(self.h, _ptr(q, _dp), len(q), int(kind), int(first_only), None, offsets.ctypes.data_as(C.POINTER(C.c_uint64)), _ptr(ids, _u32p),
               _ptr(kinds, _u8p), _ptr(pts, _dp), th)
        assert rc == 0
        tot = int(offsets[-1])
        return offsets, ids[:tot], kinds[:tot], pts[:tot]

    def first_intersection(self, rays, threads=None):
        rays = _d(rays).reshape(-1, 6)
        n = len(rays)
        prim = np.zeros(n, np.uint32)
        t = np.zeros(n)
        pts = np.zeros((n, 6))
        kind = np.zeros(n, np.uint8)
        tie = np.zeros(n, np.uint8)
        lib().orc_first_intersection(self.h, _ptr(rays, _dp), n, _ptr(prim, _u32p), _ptr(t, _dp), _ptr(pts, _dp),
                                     _ptr(kind, _u8p), _ptr(tie, _u8p), threads or nthreads())
        return dict(prim=prim, t=t, pts=pts, kind=kind, tie=tie)

    def first_intersection_skip(self, rays, skip_ids=None, skip_mask=None, threads=None):
        rays = _d(rays).reshape(-1, 6)
        n = len(rays)
        prim, t = np.zeros(n, np.uint32), np.zeros(n)
        ids = None if skip_ids is None else np.ascontiguousarray(skip_ids, dtype=np.uint32)
        words, nw = None, 0
        if skip_mask is not None:
            bits = np.packbits(np.asarray(skip_mask, dtype=bool), bitorder="little")
            bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)])
            words = np.ascontiguousarray(bits.view(np.uint32))
            nw = len(words)
        lib().orc_first_intersection_skip(self.h, _ptr(rays, _dp), n, None if ids is None else _ptr(ids, _u32p), None if words is None else _ptr(words, _u32p), nw,
                                          _ptr(prim, _u32p), _ptr(t, _dp), threads or nthreads())
        return dict(prim=prim, t=t)

    def naive_first_intersection(self, rays, threads=None):
        rays = _d(rays).reshape(-1, 6)
        n = len(rays)
        prim = np.zeros(n, np.uint32)
        t = np.zeros(n)
        ntie = np.zeros(n, np.uint32)
        lib().orc_naive_first_intersection(self.h, _ptr(rays, _dp), n, _ptr(prim, _u32p), _ptr(t, _dp), _ptr(ntie, _u32p),
                                           threads or nthreads())
        return dict(prim=prim, t=t, ntie=ntie)

    def closest_point(self, pts, threads=None):
        pts = _d(pts).reshape(-1, 3)
        n = len(pts)
        cp = np.zeros((n, 3))
        prim = np.zeros(n, np.uint32)
        sqd = np.zeros(n)
        rc = lib().orc_closest_point(self.h, _ptr(pts, _dp), n, _ptr(cp, _dp), _ptr(prim, _u32p), _ptr(sqd, _dp), threads or nthreads())
        if rc != 0:
            raise RuntimeError("CGAL_precondition(!empty()) violated")
        return dict(cp=cp, prim=prim, sqd=sqd)

    def naive_closest(self, pts, threads=None):
        pts = _d(pts).reshape(-1, 3)
        n = len(pts)
        cp = np.zeros((n, 3))
        prim = np.zeros(n, np.uint32)
        ntie = np.zeros(n, np.uint32)
        lib().orc_naive_closest(self.h, _ptr(pts, _dp), n, _ptr(cp, _dp), _ptr(prim, _u32p), _ptr(ntie, _u32p), threads or nthreads())
        return dict(cp=cp, prim=prim, ntie=ntie)

    def candidate_closest(self, face, q):
        q = _d(q)
        out = np.zeros(3)
        lib().orc_candidate_closest(self.h, int(face), _ptr(q, _dp), _ptr(out, _dp))
        return out

    def reference_point(self, face):
        """first vertex of the primitive's datum (AABB_face_graph_triangle_primitive reference point)"""
        if hasattr(self, "soup"):
            return self.soup[int(face), 0:3].copy()
        return self.verts[int(self.faces[int(face), 0])].copy()

    def candidate_ray_hit(self, face, ray):
        ray = _d(ray)
        t = C.c_double(0)
        pts = np.zeros(6)
        k = lib().orc_candidate_ray_hit(self.h, int(face), _ptr(ray, _dp), C.byref(t), _ptr(pts, _dp))
        return k, t.value, pts


# ---- predicate-level helpers
def _call(name, *arrs):
    arrs = [_d(a).ravel() for a in arrs]
    return getattr(lib(), name)(*[_ptr(a, _dp) for a in arrs])


def orientation(p, q, r, s, exact=False):
    return _call("orc_orientation_exact" if exact else "orc_orientation", np.concatenate([p, q, r, s]))


def orientation_dir(a, b, c, p, q):
    return _call("orc_orientation_dir", np.concatenate([a, b, c, p, q]))


def coplanar_orientation(*pts):
    return _call("orc_coplanar_orientation3" if len(pts) == 3 else "orc_coplanar_orientation4", np.concatenate(pts))


def compare_distance(p, q, r, exact=False):
    return _call("orc_compare_distance_exact" if exact else "orc_compare_distance", np.concatenate([p, q, r]))


def bbox_do_intersect(box, pq, is_ray, exact=False):
    box = _d(box).ravel()
    pq = _d(pq).ravel()
    f = lib().orc_bbox_do_intersect_exact if exact else lib().orc_bbox_do_intersect
    return f(_ptr(box, _dp), _ptr(pq, _dp), int(is_ray))


def classify_ray_tri(tri, pq):
    """R3T3_intersection::type of Ray_3 x Triangle_3 (-1 none, 0 facet, 1 edge, 2 vertex, 3 coplanar ray, 4 endpoint in triangle)"""
    tri, pq = _d(tri).ravel(), _d(pq).ravel()
    return lib().orc_classify_ray_tri(_ptr(tri, _dp), _ptr(pq, _dp))


def tri_do_intersect_kind(tri, q, kind):
    """kind 0 Segment_3, 1 Ray_3, 2 Line_3 (point, vector), 3 Plane_3 (a, b, c, d)"""
    tri, q = _d(tri).ravel(), _d(q).ravel()
    return lib().orc_tri_do_intersect_kind(_ptr(tri, _dp), _ptr(q, _dp), int(kind))


def bbox_do_intersect_kind(box, q, kind):
    box, q = _d(box).ravel(), _d(q).ravel()
    return lib().orc_bbox_do_intersect_kind(_ptr(box, _dp), _ptr(q, _dp), int(kind))


def tri_tri_do_intersect(query, tri):
    """Do_intersect_3(Triangle_3 query, Triangle_3 primitive)"""
    query, tri = _d(query).ravel(), _d(tri).ravel()
    return lib().orc_tri_tri_do_intersect(_ptr(query, _dp), _ptr(tri, _dp))


def tri_do_intersect(tri, pq, is_ray):
    tri = _d(tri).ravel()
    pq = _d(pq).ravel()
    return lib().orc_tri_do_intersect(_ptr(tri, _dp), _ptr(pq, _dp), int(is_ray))


def tri_query_intersection(tri, q, kind):
    """Intersect_3(Line_3 (kind 2: point, vector) | Plane_3 (kind 3: a, b, c, d), Triangle_3): (object kind 0..3, points[2,3])"""
    tri, q = _d(tri).ravel(), _d(q).ravel()
    out = np.zeros(6)
    k = lib().orc_tri_query_intersection(_ptr(tri, _dp), _ptr(q, _dp), int(kind), _ptr(out, _dp))
    return k, out.reshape(2, 3)


def tri_seg_intersection(tri, seg):
    tri, seg = _d(tri).ravel(), _d(seg).ravel()
    out = np.zeros(6)
    k = lib().orc_tri_seg_intersection(_ptr(tri, _dp), _ptr(seg, _dp), _ptr(out, _dp))
    return k, out.reshape(2, 3)


def tri_ray_intersection(tri, ray):
    tri = _d(tri).ravel()
    ray = _d(ray).ravel()
    out = np.zeros(6)
    k = lib().orc_tri_ray_intersection(_ptr(tri, _dp), _ptr(ray, _dp), _ptr(out, _dp))
    return k, out.reshape(2, 3)


def project_on_triangle(tri, q):
    tri = _d(tri).ravel()
    q = _d(q).ravel()
    out = np.zeros(3)
    lib().orc_project_on_triangle(_ptr(tri, _dp), _ptr(q, _dp), _ptr(out, _dp))
    return out


def ball_reaches_box(q, box, bound):
    q, box, bound = _d(q).ravel(), _d(box).ravel(), _d(bound).ravel()
    return lib().orc_ball_reaches_box(_ptr(q, _dp), _ptr(box, _dp), _ptr(bound, _dp))


def intersection_distance(ray, box):
    ray, box = _d(ray).ravel(), _d(box).ravel()
    t = C.c_double(0)
    ok = lib().orc_intersection_distance(_ptr(ray, _dp), _ptr(box, _dp), C.byref(t))
    return (t.value if ok else None)
