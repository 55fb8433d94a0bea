"""ctypes binding of tests/hostcheck/libhostcheck.so: the device headers compiled for the host (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def warp_bounds(tri, q, mag):
    """(lower1, lower2, upper or nan, e2, dslack) of closest_warp.cuh's pre-tests for one (query, primitive) pair"""
    t, p, out = np.ascontiguousarray(tri, np.float64), np.ascontiguousarray(q, np.float64), np.zeros(5, np.float32)
    lib().hc_warp_bounds(t.ctypes.data_as(_dp), p.ctypes.data_as(_dp), float(mag), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def build():
    d = os.path.join(ROOT, "tests", "hostcheck")
    src, out = os.path.join(d, "hostcheck.cu"), os.path.join(d, "libhostcheck.so")
    deps = [src] + [os.path.join(ROOT, "cgal_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "cgal_b200", "csrc")) if f.endswith(".cuh")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(p) for p in deps):
        return out
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared",
                           "-Xcompiler", "-fPIC,-ffp-contract=off", "-o", out, src])
    return out


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.hc_build.restype = C.c_void_p
        L.hc_build.argtypes = [_dp, _u32p, C.c_uint64]
        L.hc_destroy.argtypes = [C.c_void_p]
        L.hc_warp_bounds.argtypes = [_dp, _dp, C.c_double, C.POINTER(C.c_float)]
        L.hc_num_nodes.restype = C.c_uint64
        L.hc_num_nodes.argtypes = [C.c_void_p]
        L.hc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _dp]
        L.hc_validate.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.hc_intersect.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, C.c_int, _u32p]
        L.hc_first_intersection.argtypes = [C.c_void_p, _dp, C.c_uint64, _u32p, _dp, _dp, _u8p, _u8p]
        L.hc_closest_point.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp, _u32p, _dp, _u8p]
        L.hc_closest_point_segments.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp, _u32p, _dp, _u8p]
        L.hc_count_triangle_segments.argtypes = [C.c_void_p, _dp, C.c_uint64, _u32p]
        L.hc_closest_point_pair.argtypes = [C.c_void_p, _dp, C.c_uint64, _u32p, C.c_int, C.c_int, _dp, _u32p, _dp, _u8p, C.POINTER(C.c_uint64)]
        for n in ("hc_orient3", "hc_orient3_exact", "hc_orient3_dir", "hc_coplanar_orient", "hc_cmp_dist", "hc_cmp_dist_exact"):
            getattr(L, n).argtypes = [_dp]
        L.hc_tri_do_intersect.argtypes = [_dp, _dp, C.c_int]
        L.hc_tri_tri_do_intersect.argtypes = [_dp, _dp]
        L.hc_side_of_mesh_xf.argtypes = [C.c_void_p, _dp, _dp, C.c_uint64, _dp, C.c_int, C.POINTER(C.c_int8)]
        L.hc_parse_decimals.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.hc_tree_tree_xf.argtypes = [C.c_void_p, _dp, C.c_void_p, _dp, _u32p]
        L.hc_xf_apply.argtypes = [_dp, _dp, _dp]
        L.hc_xf_box_meets.argtypes = [_dp, C.POINTER(C.c_float), _dp]
        L.hc_intersect_triangle.argtypes = [C.c_void_p, _dp, C.c_uint64, C.c_int, _u32p]
        L.hc_tri_ray_intersection.argtypes = [_dp, _dp, _dp, _dp]
        L.hc_closest_on_triangle.argtypes = [_dp, _dp, _dp]
        L.hc_line_box.argtypes = [_dp, _dp, C.c_int]
        L.hc_point_box_sqdist.argtypes = [_dp, _dp]
        L.hc_point_box_sqdist.restype = C.c_double
        L.hc_sizeof_node.restype = C.c_uint64
        L.hc_sizeof_tri.restype = C.c_uint64
        _LIB = L
    return _LIB


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, ty=_dp):
    return None if a is None else a.ctypes.data_as(ty)


class HostTree:
    def __init__(self, soup, ids=None):
        self.soup = _d(soup).reshape(-1, 9)
        self.ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32)
        self.n = len(self.soup)
        self.h = lib().hc_build(_p(self.soup), _p(self.ids, _u32p), self.n)

    def __del__(self):
        if getattr(self, "h", None):
            lib().hc_destroy(self.h)
            self.h = None

    def export(self):
        nn = lib().hc_num_nodes(self.h)
        nodes = np.zeros(nn * 64, np.uint8)
        tris = np.zeros(self.n * 96, np.uint8)
        bbox = np.zeros(6)
        lib().hc_export(self.h, nodes.ctypes.data_as(C.c_void_p), tris.ctypes.data_as(C.c_void_p), _p(bbox))
        return nodes, tris, bbox

    def fused_build_diff(self, frontier_per_block=-1):
        """(number of node words in which the fused bottom-up build differs from the classic one, frontier entries, nodes left to the searches, overflow)"""
        st = np.zeros(3, np.uint64)
        f = lib().hc_fused_build_diff
        f.restype = C.c_int64
        d = f(self.h, int(frontier_per_block), st.ctypes.data_as(C.POINTER(C.c_uint64)))
        return int(d), int(st[0]), int(st[1]), int(st[2])

    def intersect(self, q, is_ray, count):
        q = _d(q).reshape(-1, 6)
        out = np.zeros(len(q), np.uint32)
        lib().hc_intersect(self.h, _p(q), len(q), int(is_ray), int(count), _p(out, _u32p))
        return out

    def intersect_triangle(self, q, count):
        """do_intersect / number_of_intersected_primitives of Triangle_3 queries (9 doubles) through bvh.cuh::traverse_triangle"""
        q = _d(q).reshape(-1, 9)
        out = np.zeros(len(q), np.uint32)
        lib().hc_intersect_triangle(self.h, _p(q), len(q), int(count), _p(out, _u32p))
        return out

    def count_line_plane(self, q, kind):
        """number_of_intersected_primitives of Line_3 (kind 2: point + vector) / Plane_3 (kind 3: a, b, c, d) queries through bvh.cuh"""
        q = _d(q).reshape(-1, 4 if kind == 3 else 6)
        out = np.zeros(len(q), np.uint32)
        lib().hc_count_line_plane(self.h, _p(q), len(q), int(kind), _p(out, _u32p))
        return out

    def side_of_mesh(self, pts, dirs):
        """Side_of_triangle_mesh through bvh.cuh::traverse_side with the given ray directions; (side, rays cast); side -2 = no decisive ray"""
        pts, dirs = _d(pts).reshape(-1, 3), _d(dirs).reshape(-1, 3)
        side, att = np.zeros(len(pts), np.int8), np.zeros(len(pts), np.uint8)
        lib().hc_side_of_mesh(self.h, _p(pts), len(pts), _p(dirs), len(dirs), side.ctypes.data_as(C.POINTER(C.c_int8)), _p(att, _u8p))
        return side, att

    def side_of_mesh_xf(self, m, pts, dirs):
        """the same with the mesh under the 3 x 4 affine map m (bvh.cuh::traverse_side<AffXf>): side per point, -2 = no decisive ray"""
        m, pts, dirs = _d(m).ravel(), _d(pts).reshape(-1, 3), _d(dirs).reshape(-1, 3)
        side = np.zeros(len(pts), np.int8)
        lib().hc_side_of_mesh_xf(self.h, _p(m), _p(pts), len(pts), _p(dirs), len(dirs), side.ctypes.data_as(C.POINTER(C.c_int8)))
        return side

    def do_intersect_tree_xf(self, m, other, other_m):
        """tree-vs-tree under affine maps as the device kernel runs it (this tree is traversed, `other`'s primitives are the queries): (bool, id here, id there)"""
        pair = np.full(2, 0xFFFFFFFF, np.uint32)
        r = lib().hc_tree_tree_xf(self.h, _p(_d(m).ravel()), other.h, _p(_d(other_m).ravel()), _p(pair, _u32p))
        return bool(r), int(pair[1]), int(pair[0])

    def first_intersection(self, rays):
        rays = _d(rays).reshape(-1, 6)
        n = len(rays)
        prim, t, pts = np.zeros(n, np.uint32), np.zeros(n), np.zeros((n, 6))
        kind, tie = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        lib().hc_first_intersection(self.h, _p(rays), n, _p(prim, _u32p), _p(t), _p(pts), _p(kind, _u8p), _p(tie, _u8p))
        return dict(prim=prim, t=t, pts=pts, kind=kind, tie=tie & 1, near=(tie >> 1) & 1)

    def closest_point(self, q):
        q = _d(q).reshape(-1, 3)
        n = len(q)
        cp, prim, sqd, tie = np.zeros((n, 3)), np.zeros(n, np.uint32), np.zeros(n), np.zeros(n, np.uint8)
        lib().hc_closest_point(self.h, _p(q), n, _p(cp), _p(prim, _u32p), _p(sqd), _p(tie, _u8p))
        return dict(cp=cp, prim=prim, sqd=sqd, tie=tie & 1, near=(tie >> 1) & 1)


    def closest_point_segments(self, q):
        """closest_point on a tree of Segment_3 primitives stored as (source, target, target) records (bvh.cuh::traverse_closest<true>)"""
        q = _d(q).reshape(-1, 3)
        n = len(q)
        cp, prim, sqd, tie = np.zeros((n, 3)), np.zeros(n, np.uint32), np.zeros(n), np.zeros(n, np.uint8)
        lib().hc_closest_point_segments(self.h, _p(q), n, _p(cp), _p(prim, _u32p), _p(sqd), _p(tie, _u8p))
        return dict(cp=cp, prim=prim, sqd=sqd, tie=tie & 1, near=(tie >> 1) & 1)

    def count_triangle_segments(self, q):
        """number_of_intersected_primitives(Triangle_3) on such a tree (bvh.cuh::traverse_triangle<kCount, NoXf, true>)"""
        q = _d(q).reshape(-1, 9)
        out = np.zeros(len(q), np.uint32)
        lib().hc_count_triangle_segments(self.h, _p(q), len(q), _p(out, _u32p))
        return out

    def closest_point_pair(self, q, perm=None, warps=1, exact_go=64):
        """closest_pair.cuh's kernel body on emulated warps (32 fibers each): the code the GPU runs"""
        q = _d(q).reshape(-1, 3)
        n = len(q)
        cp, prim, sqd, fl = np.zeros((n, 3)), np.full(n, 0xFFFFFFFF, np.uint32), np.zeros(n), np.zeros(n, np.uint8)
        stats = np.zeros(3, np.uint64)
        pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.uint32)
        rc = lib().hc_closest_point_pair(self.h, _p(q), n, _p(pm, _u32p), int(warps), int(exact_go), _p(cp), _p(prim, _u32p), _p(sqd), _p(fl, _u8p),
                                         stats.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert rc == 0, "emulated pair kernel: %s" % {1: "dead-lock (non-uniform control flow at a collective)", 2: "overflow reported"}.get(rc, rc)
        return dict(cp=cp, prim=prim, sqd=sqd, tie=(fl >> 2) & 1, exact=(fl >> 3) & 1, nodes=int(stats[0]), leaves=int(stats[1]), culled=int(stats[2]))


def validate(nodes_u8, tris_u8, n):
    return lib().hc_validate(nodes_u8.ctypes.data_as(C.c_void_p), tris_u8.ctypes.data_as(C.c_void_p), n)


def call(name, *arrs):
    arrs = [_d(a).ravel() for a in arrs]
    return getattr(lib(), name)(*[_p(a) for a in arrs])


def classify_ray_tri(tri, pq):
    tri, pq = _d(tri).ravel(), _d(pq).ravel()
    return lib().hc_classify_ray_tri(_p(tri), _p(pq))


def tri_do_intersect_kind(tri, q, kind):
    tri, q = _d(tri).ravel(), _d(q).ravel()
    return lib().hc_tri_do_intersect_kind(_p(tri), _p(q), int(kind))


def tri_tri_do_intersect(query, tri):
    query, tri = _d(query).ravel(), _d(tri).ravel()
    return lib().hc_tri_tri_do_intersect(_p(query), _p(tri))


def parse_decimals(literals):
    """ingest.cuh::parse_decimal on every string: (values as float64 from the returned bit patterns, status per literal: 0 ok, 1 not a number, 2 left to strtod)"""
    enc = [x.encode() for x in literals]
    offs = np.zeros(len(enc) + 1, np.uint64)
    np.cumsum([len(e) for e in enc], out=offs[1:])
    buf = b"".join(enc) + b" "
    bits, status = np.zeros(len(enc), np.uint64), np.zeros(len(enc), np.int32)
    lib().hc_parse_decimals(buf, offs.ctypes.data_as(C.POINTER(C.c_uint64)), len(enc), bits.ctypes.data_as(C.POINTER(C.c_uint64)), status.ctypes.data_as(C.POINTER(C.c_int)))
    return bits.view(np.float64), status


def xf_apply(m, tri):
    """bvh.cuh::AffXf::apply: the triangle's image under the 3 x 4 matrix m (12 doubles, row major)"""
    m, tri, out = _d(m).ravel(), _d(tri).ravel(), np.zeros(9)
    lib().hc_xf_apply(_p(m), _p(tri), _p(out))
    return out


def xf_box_meets(m, box_lo, box_hi, query):
    """bvh.cuh::AffXf::box_meets: may the image of the binary32 box [lo, hi] meet the (binary32, outward) box of the query triangle?"""
    m, query = _d(m).ravel(), _d(query).ravel()
    b = np.concatenate([np.asarray(box_lo, np.float32), np.asarray(box_hi, np.float32)])
    return bool(lib().hc_xf_box_meets(_p(m), b.ctypes.data_as(C.POINTER(C.c_float)), _p(query)))


def tri_do_intersect(tri, pq, is_ray):
    tri, pq = _d(tri).ravel(), _d(pq).ravel()
    return lib().hc_tri_do_intersect(_p(tri), _p(pq), int(is_ray))


def tri_query_intersection(tri, q, kind):
    tri, q = _d(tri).ravel(), _d(q).ravel()
    pts = np.zeros(6)
    k = lib().hc_tri_query_intersection(_p(tri), _p(q), int(kind), _p(pts))
    return k, pts.reshape(2, 3)


def tri_seg_intersection(tri, seg):
    tri, seg = _d(tri).ravel(), _d(seg).ravel()
    pts = np.zeros(6)
    k = lib().hc_tri_seg_intersection(_p(tri), _p(seg), _p(pts))
    return k, pts.reshape(2, 3)


def tri_ray_intersection(tri, ray):
    tri, ray = _d(tri).ravel(), _d(ray).ravel()
    pts = np.zeros(6)
    t = C.c_double(0)
    k = lib().hc_tri_ray_intersection(_p(tri), _p(ray), _p(pts), C.byref(t))
    return k, pts.reshape(2, 3), t.value


def closest_on_triangle(tri, q):
    tri, q = _d(tri).ravel(), _d(q).ravel()
    out = np.zeros(3)
    lib().hc_closest_on_triangle(_p(tri), _p(q), _p(out))
    return out


def point_box_sqdist(box, p):
    box, p = _d(box).ravel(), _d(p).ravel()
    return lib().hc_point_box_sqdist(_p(box), _p(p))


def line_box(box, pq, is_ray):
    box, pq = _d(box).ravel(), _d(pq).ravel()
    return lib().hc_line_box(_p(box), _p(pq), int(is_ray))
