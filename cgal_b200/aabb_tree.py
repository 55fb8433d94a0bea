"""Host-side mirror of CGAL::AABB_tree (AABB_tree/include/CGAL/AABB_tree.h:55-677) over the C ABI of libb200aabb.so.

Same member names and meaning as the reference class, with batched arguments: a query argument is an (n, 6) array of
Segment_3 / Ray_3 records (source, target|second point) or an (n, 3) array of Point_3; results are arrays.
numpy arrays are HOST buffers (copied inside the call); torch CUDA tensors are used in place (DEVICE pointers).
"""
import ctypes as C
import os

import numpy as np

SEGMENT, RAY, LINE, PLANE, TRIANGLE = 0, 1, 2, 3, 4      # LINE: (point, direction vector) per query; PLANE: (a, b, c, d); TRIANGLE: three vertices (id queries only)
_QUERY_DOUBLES = {PLANE: 4, TRIANGLE: 9}                  # doubles per query of a kind (every other kind: 6)
FLAG_HIT, FLAG_SEGMENT, FLAG_TIE, FLAG_EXACT, FLAG_NEAR_TIE, FLAG_TRIANGLE = 1, 2, 4, 8, 16, 32
OPT_SORT_QUERIES = 1
OPT_CLOSEST_KERNEL = 2    # 1 = warp-cooperative work pool (default), 2 = lane pairs with private stacks, 3 = two passes (binary32 filter + exact resolve), 0 = one thread per query
OPT_WARP_EXACT_GO = 4
OPT_HOST_PARTS = 5
OPT_ASYNC_HOST = 6
OPT_RAY_KERNEL = 7
NO_PRIMITIVE = 0xFFFFFFFF
_HOST, _DEVICE = 0, 1
_ERRORS = {-1: "invalid argument", -2: "CUDA error", -3: "out of device memory", -4: "empty tree (CGAL_precondition(!empty()))",
           -5: "no sm_100 CUDA device (there is no CPU fallback)", -6: "traversal incomplete", -7: "NCCL unavailable or failed"}
_LIB = None


class B200AABBError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__("b200aabb: %s (%d)%s" % (_ERRORS.get(code, "error"), code, (": " + detail) if detail else ""))


class _Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("exact_fallbacks", C.c_uint64), ("ties", C.c_uint64),
                ("near_ties", C.c_uint64), ("nodes_visited", C.c_uint64), ("leaves_tested", C.c_uint64), ("leaves_culled", C.c_uint64),
                ("last_build_ms", C.c_double), ("last_query_ms", C.c_double), ("last_query_total_ms", C.c_double)]


SYMBOLS = ["b200aabb_create", "b200aabb_destroy", "b200aabb_set_stream", "b200aabb_set_option", "b200aabb_last_error", "b200aabb_get_stats", "b200aabb_synchronize",
           "b200aabb_set_mesh", "b200aabb_set_mesh_off", "b200aabb_set_triangles", "b200aabb_set_segments", "b200aabb_clear", "b200aabb_build", "b200aabb_refit", "b200aabb_refit_mesh",
           "b200aabb_accelerate_distance_queries", "b200aabb_size", "b200aabb_bbox", "b200aabb_do_intersect",
           "b200aabb_number_of_intersected_primitives", "b200aabb_any_intersected_primitive", "b200aabb_all_intersected_primitives", "b200aabb_do_intersect_tree", "b200aabb_do_intersect_tree_transformed",
           "b200aabb_number_of_intersections", "b200aabb_all_intersections", "b200aabb_any_intersection",
           "b200aabb_number_of_intersections_q", "b200aabb_all_intersections_q", "b200aabb_any_intersection_q",
           "b200aabb_first_intersection", "b200aabb_first_intersection_skip", "b200aabb_closest_point", "b200aabb_max_distance_to_mesh", "b200aabb_side_of_mesh", "b200aabb_side_of_mesh_transformed",
           "b200aabb_export_layout", "b200aabb_import_alloc", "b200aabb_import_finish", "b200aabb_debug_predicate", "b200aabb_version",
           "b200aabb_comm_unique_id", "b200aabb_comm_init", "b200aabb_comm_init_all", "b200aabb_comm_destroy", "b200aabb_comm_last_error",
           "b200aabb_comm_broadcast_tree", "b200aabb_comm_broadcast_tree_all"]


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200aabb.so")


def load_library():
    """Loads libb200aabb.so.  Raises if it has not been built: the CUDA library IS the implementation."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError("libb200aabb.so is missing (%s): run `python -m cgal_b200._build`; there is no fallback path" % path)
        L = C.CDLL(path)
        vp, u64, dp = C.c_void_p, C.c_uint64, C.c_void_p
        L.b200aabb_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.b200aabb_destroy.argtypes = [vp]
        L.b200aabb_set_stream.argtypes = [vp, vp]
        L.b200aabb_set_option.argtypes = [vp, C.c_int, C.c_int64]
        L.b200aabb_last_error.argtypes = [vp]
        L.b200aabb_last_error.restype = C.c_char_p
        L.b200aabb_get_stats.argtypes = [vp, C.POINTER(_Stats)]
        L.b200aabb_synchronize.argtypes = [vp]
        L.b200aabb_set_mesh.argtypes = [vp, dp, u64, dp, dp, u64]
        L.b200aabb_set_triangles.argtypes = [vp, dp, dp, u64, C.c_int]
        L.b200aabb_set_segments.argtypes = [vp, dp, dp, u64, C.c_int]
        L.b200aabb_set_mesh_off.argtypes = [vp, C.c_char_p, u64]
        L.b200aabb_clear.argtypes = [vp]
        L.b200aabb_refit.argtypes = [vp, dp, u64, C.c_int]
        L.b200aabb_refit_mesh.argtypes = [vp, dp, u64, dp, u64]
        L.b200aabb_build.argtypes = [vp]
        L.b200aabb_accelerate_distance_queries.argtypes = [vp]
        L.b200aabb_size.argtypes = [vp, C.POINTER(u64)]
        L.b200aabb_bbox.argtypes = [vp, C.POINTER(C.c_double)]
        L.b200aabb_do_intersect.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp]
        L.b200aabb_number_of_intersected_primitives.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp]
        L.b200aabb_any_intersected_primitive.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp]
        L.b200aabb_all_intersected_primitives.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp, dp]
        L.b200aabb_do_intersect_tree.argtypes = [vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.b200aabb_do_intersect_tree_transformed.argtypes = [vp, dp, vp, dp, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.b200aabb_number_of_intersections.argtypes = [vp, dp, u64, C.c_int, dp]
        L.b200aabb_all_intersections.argtypes = [vp, dp, u64, C.c_int, dp, dp, dp, dp]
        L.b200aabb_any_intersection.argtypes = [vp, dp, u64, C.c_int, dp, dp, dp, dp]
        L.b200aabb_number_of_intersections_q.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp]
        L.b200aabb_all_intersections_q.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp, dp, dp, dp]
        L.b200aabb_any_intersection_q.argtypes = [vp, dp, u64, C.c_int, C.c_int, dp, dp, dp, dp]
        L.b200aabb_first_intersection.argtypes = [vp, dp, u64, C.c_int, dp, dp, dp, dp, dp]
        L.b200aabb_first_intersection_skip.argtypes = [vp, dp, u64, C.c_int, dp, dp, C.c_uint32, dp, dp, dp, dp, dp]
        L.b200aabb_closest_point.argtypes = [vp, dp, u64, C.c_int, dp, dp, dp, dp]
        L.b200aabb_max_distance_to_mesh.argtypes = [vp, dp, u64, C.c_int, C.POINTER(C.c_double), C.POINTER(u64)]
        L.b200aabb_side_of_mesh.argtypes = [vp, dp, u64, C.c_int, dp, dp]
        L.b200aabb_side_of_mesh_transformed.argtypes = [vp, dp, dp, u64, C.c_int, dp, dp]
        L.b200aabb_export_layout.argtypes = [vp, C.POINTER(vp), C.POINTER(u64), C.c_int, C.POINTER(C.c_int)]
        L.b200aabb_import_alloc.argtypes = [vp, u64]
        L.b200aabb_import_finish.argtypes = [vp]
        L.b200aabb_debug_predicate.argtypes = [C.c_int, C.c_int, dp, C.c_int, u64, dp, dp, dp]
        L.b200aabb_version.restype = C.c_char_p
        L.b200aabb_comm_unique_id.argtypes = [C.c_char_p]
        L.b200aabb_comm_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(vp)]
        L.b200aabb_comm_init_all.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
        L.b200aabb_comm_destroy.argtypes = [vp]
        L.b200aabb_comm_last_error.argtypes = [vp]
        L.b200aabb_comm_last_error.restype = C.c_char_p
        L.b200aabb_comm_broadcast_tree.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(u64)]
        L.b200aabb_comm_broadcast_tree_all.argtypes = [C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(u64)]
        _LIB = L
    return _LIB


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Buf:
    """a query / result buffer: numpy (host) or torch CUDA tensor (device)"""

    def __init__(self, arr, dtype, cols):
        if _is_torch(arr):
            import torch
            want = {np.float64: torch.float64, np.uint32: torch.int32, np.uint8: torch.uint8, np.int8: torch.int8}[dtype]
            if not arr.is_cuda:
                raise ValueError("torch tensors must live on the GPU; pass numpy arrays for host buffers")
            if arr.dtype != want or not arr.is_contiguous():
                raise ValueError("expected a contiguous %s CUDA tensor" % want)
            self.device, self.obj, self.ptr, self.n = True, arr, arr.data_ptr(), (arr.numel() // cols)
        else:
            a = np.ascontiguousarray(arr, dtype=dtype)
            self.device, self.obj, self.ptr, self.n = False, a, a.ctypes.data, (a.size // cols)


def _alloc_like(q, n, dtype, cols):
    if q.device:
        import torch
        tdt = {np.float64: torch.float64, np.uint32: torch.int32, np.uint8: torch.uint8, np.int8: torch.int8}[dtype]
        shape = (n, cols) if cols > 1 else (n,)
        return _Buf(torch.empty(shape, dtype=tdt, device=q.obj.device), dtype, cols)
    shape = (n, cols) if cols > 1 else (n,)
    return _Buf(np.empty(shape, dtype=dtype), dtype, cols)


def debug_predicate(which, records, device=0):
    """self-test hook: one device predicate per row of `records`; returns (result int32[n], used_exact int32[n], values f64[n,8])"""
    L = load_library()
    r = np.ascontiguousarray(records, dtype=np.float64)
    r = r.reshape(len(r), -1)
    n = len(r)
    res, ex, val = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros((n, 8))
    rc = L.b200aabb_debug_predicate(int(device), int(which), r.ctypes.data, r.shape[1], n, res.ctypes.data, ex.ctypes.data, val.ctypes.data)
    if rc:
        raise B200AABBError(rc)
    return res, ex, val


class AABB_tree:
    """AABB_tree< AABB_traits_3< Epick, AABB_face_graph_triangle_primitive< Surface_mesh<Point_3> > > > on one B200."""

    def __init__(self, vertices=None, faces=None, triangles=None, ids=None, device=0, segments=None):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.b200aabb_create(int(device), C.byref(h))
        if rc:
            raise B200AABBError(rc)
        self._h = h
        self.device = int(device)
        if vertices is not None or triangles is not None or segments is not None:
            self.insert(vertices=vertices, faces=faces, triangles=triangles, ids=ids, segments=segments)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b200aabb_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc:
            raise B200AABBError(rc, self._lib.b200aabb_last_error(self._h).decode())

    # ---- primitives (AABB_tree.h:147-216) ----
    def insert(self, vertices=None, faces=None, triangles=None, ids=None, segments=None):
        """insert(faces(mesh).first, faces(mesh).second, mesh): `vertices` (nv,3) f64 + `faces` (nf,3) u32 in face order;
        or a triangle soup `triangles` (n,9); or Segment_3 primitives `segments` (n,6) = (source, target) (AABB_segment_primitive_3 /
        AABB_halfedge_graph_segment_primitive: closest_point and Plane_3 / Triangle_3 id queries).  Replaces the current primitives."""
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            idp = ids.ctypes.data
        if segments is not None:
            if _is_torch(segments):
                b = _Buf(segments, np.float64, 6)
                if ids is not None:
                    raise ValueError("ids with device segments are not supported")
                self._check(self._lib.b200aabb_set_segments(self._h, b.ptr, None, b.n, _DEVICE))
            else:
                sg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 6)
                self._check(self._lib.b200aabb_set_segments(self._h, sg.ctypes.data, idp, len(sg), _HOST))
            return
        if triangles is not None:
            if _is_torch(triangles):
                b = _Buf(triangles, np.float64, 9)
                if ids is not None:
                    raise ValueError("ids with device triangles must be passed as a CUDA tensor via insert_device")
                self._check(self._lib.b200aabb_set_triangles(self._h, b.ptr, None, b.n, _DEVICE))
            else:
                t = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 9)
                self._check(self._lib.b200aabb_set_triangles(self._h, t.ctypes.data, idp, len(t), _HOST))
        else:
            v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
            f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
            self._check(self._lib.b200aabb_set_mesh(self._h, v.ctypes.data, len(v), f.ctypes.data, idp, len(f)))

    def insert_off(self, text):
        """the mesh as the text (bytes) of a plain triangle OFF file, tokenised and parsed on the device (b200aabb_set_mesh_off)"""
        data = text if isinstance(text, (bytes, bytearray)) else str(text).encode()
        self._check(self._lib.b200aabb_set_mesh_off(self._h, bytes(data), len(data)))

    def clear(self):
        self._check(self._lib.b200aabb_clear(self._h))

    def build(self):
        self._check(self._lib.b200aabb_build(self._h))

    rebuild = build

    def refit(self, vertices=None, faces=None, triangles=None):
        """the primitives have moved (same count and order): new coordinates, boxes recomputed over the existing hierarchy (b200aabb_refit)"""
        if triangles is not None:
            if _is_torch(triangles):
                b = _Buf(triangles, np.float64, 9)
                self._check(self._lib.b200aabb_refit(self._h, b.ptr, b.n, _DEVICE))
            else:
                t = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 9)
                self._check(self._lib.b200aabb_refit(self._h, t.ctypes.data, len(t), _HOST))
        else:
            v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
            f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
            self._check(self._lib.b200aabb_refit_mesh(self._h, v.ctypes.data, len(v), f.ctypes.data, len(f)))

    def accelerate_distance_queries(self):
        self._check(self._lib.b200aabb_accelerate_distance_queries(self._h))
        return True

    def do_not_accelerate_distance_queries(self):
        pass   # the BVH-derived bound needs no auxiliary structure

    def size(self):
        n = C.c_uint64()
        self._check(self._lib.b200aabb_size(self._h, C.byref(n)))
        return n.value

    def empty(self):
        return self.size() == 0

    def bbox(self):
        out = (C.c_double * 6)()
        self._check(self._lib.b200aabb_bbox(self._h, out))
        return np.array(out[:])

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.b200aabb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, option, value):
        self._check(self._lib.b200aabb_set_option(self._h, int(option), int(value)))

    def synchronize(self):
        """waits for the batches in flight (OPT_ASYNC_HOST); raises if one of them could not finish"""
        self._check(self._lib.b200aabb_synchronize(self._h))

    def stats(self):
        s = _Stats()
        self._check(self._lib.b200aabb_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _Stats._fields_}

    # ---- intersection queries (AABB_tree.h:243-354) ----
    def do_intersect(self, queries, kind=SEGMENT, out=None):
        q = _Buf(queries, np.float64, _QUERY_DOUBLES.get(kind, 6))
        o = _Buf(out, np.uint8, 1) if out is not None else _alloc_like(q, q.n, np.uint8, 1)
        self._check(self._lib.b200aabb_do_intersect(self._h, q.ptr, q.n, kind, int(q.device), o.ptr))
        return o.obj

    def number_of_intersected_primitives(self, queries, kind=SEGMENT, out=None):
        q = _Buf(queries, np.float64, _QUERY_DOUBLES.get(kind, 6))
        o = _Buf(out, np.uint32, 1) if out is not None else _alloc_like(q, q.n, np.uint32, 1)
        self._check(self._lib.b200aabb_number_of_intersected_primitives(self._h, q.ptr, q.n, kind, int(q.device), o.ptr))
        return o.obj

    def any_intersected_primitive(self, queries, kind=SEGMENT, out=None):
        """any_intersected_primitive(query): id of one intersected primitive per query, NO_PRIMITIVE = nullopt"""
        q = _Buf(queries, np.float64, _QUERY_DOUBLES.get(kind, 6))
        o = _Buf(out, np.uint32, 1) if out is not None else _alloc_like(q, q.n, np.uint32, 1)
        self._check(self._lib.b200aabb_any_intersected_primitive(self._h, q.ptr, q.n, kind, int(q.device), o.ptr))
        return o.obj

    def all_intersected_primitives(self, queries, kind=SEGMENT):
        """all_intersected_primitives(query, out) for a HOST batch: (offsets uint64[n+1], ids uint32[total]); the ids of query i are
        ids[offsets[i]:offsets[i+1]] (count pass -> prefix sums -> fill pass, include/b200aabb.h)"""
        q = _Buf(np.asarray(queries), np.float64, _QUERY_DOUBLES.get(kind, 6))
        counts = self.number_of_intersected_primitives(q.obj, kind)
        offsets = np.zeros(q.n + 1, np.uint64)
        np.cumsum(counts, out=offsets[1:])
        ids = np.empty(max(int(offsets[-1]), 1), np.uint32)
        self._check(self._lib.b200aabb_all_intersected_primitives(self._h, q.ptr, q.n, kind, 0, offsets.ctypes.data, ids.ctypes.data))
        return offsets, ids[:int(offsets[-1])]

    def do_intersect_tree(self, other, witness=False, xf=None, other_xf=None):
        """tree.do_intersect(other_tree) (AABB_traits_3.h:679-690): True iff a primitive of this tree intersects a primitive of `other` (same device);
        witness=True -> (bool, id in this tree, id in other) of one intersecting pair (NO_PRIMITIVE twice when there is none).
        xf / other_xf: 3 x 4 affine maps (Aff_transformation_3::m(i, j), row major) the two meshes stand under, as in
        Rigid_triangle_mesh_collision_detection::does_A_intersect_B; None = identity"""
        r, a, b = C.c_int(0), C.c_uint32(NO_PRIMITIVE), C.c_uint32(NO_PRIMITIVE)
        if xf is not None or other_xf is not None:
            m1 = None if xf is None else np.ascontiguousarray(xf, np.float64).reshape(12)
            m2 = None if other_xf is None else np.ascontiguousarray(other_xf, np.float64).reshape(12)
            self._check(self._lib.b200aabb_do_intersect_tree_transformed(self._h, None if m1 is None else m1.ctypes.data, other._h, None if m2 is None else m2.ctypes.data,
                                                                         C.byref(r), C.byref(a), C.byref(b)))
            return (bool(r.value), a.value, b.value) if witness else bool(r.value)
        self._check(self._lib.b200aabb_do_intersect_tree(self._h, other._h, C.byref(r), C.byref(a), C.byref(b)))
        return (bool(r.value), a.value, b.value) if witness else bool(r.value)

    def all_intersections(self, rays, kind=RAY):
        """all_intersections(Ray_3 | Segment_3 | Line_3 | Plane_3, out) for a HOST batch: (offsets uint64[n+1], ids uint32[total], kinds uint8[total] (1 Point_3 /
        2 Segment_3 / 3 the primitive's Triangle_3), points f64[total, 6]); the entries of query i are [offsets[i], offsets[i+1])"""
        q = _Buf(np.asarray(rays), np.float64, _QUERY_DOUBLES.get(kind, 6))
        counts = np.empty(q.n, np.uint32)
        self._check(self._lib.b200aabb_number_of_intersections_q(self._h, q.ptr, q.n, int(kind), 0, counts.ctypes.data))
        offsets = np.zeros(q.n + 1, np.uint64)
        np.cumsum(counts, out=offsets[1:])
        tot = max(int(offsets[-1]), 1)
        ids, kinds, pts = np.empty(tot, np.uint32), np.empty(tot, np.uint8), np.empty((tot, 6))
        self._check(self._lib.b200aabb_all_intersections_q(self._h, q.ptr, q.n, int(kind), 0, offsets.ctypes.data, ids.ctypes.data, kinds.ctypes.data, pts.ctypes.data))
        tot = int(offsets[-1])
        return offsets, ids[:tot], kinds[:tot], pts[:tot]

    def any_intersection(self, rays, want=("prim", "point", "point2", "flags"), out=None, kind=RAY):
        """any_intersection(Ray_3 | Segment_3 | Line_3 | Plane_3): dict with prim (NO_PRIMITIVE = nullopt), point, point2, flags (FLAG_HIT, FLAG_SEGMENT,
        FLAG_TRIANGLE)"""
        q = _Buf(rays, np.float64, _QUERY_DOUBLES.get(kind, 6))
        spec = {"prim": (np.uint32, 1), "point": (np.float64, 3), "point2": (np.float64, 3), "flags": (np.uint8, 1)}
        bufs = {}
        for k in want:
            bufs[k] = _Buf(out[k], *spec[k]) if out is not None and k in out else _alloc_like(q, q.n, *spec[k])
        p = lambda k: bufs[k].ptr if k in bufs else None   # noqa: E731
        self._check(self._lib.b200aabb_any_intersection_q(self._h, q.ptr, q.n, int(kind), int(q.device), p("prim"), p("point"), p("point2"), p("flags")))
        return {k: b.obj for k, b in bufs.items()}

    def first_intersection(self, rays, want=("prim", "t", "point", "point2", "flags"), out=None):
        """first_intersection(Ray_3) for every ray: dict with prim (NO_PRIMITIVE = nullopt), t, point, point2, flags."""
        q = _Buf(rays, np.float64, 6)
        spec = {"prim": (np.uint32, 1), "t": (np.float64, 1), "point": (np.float64, 3), "point2": (np.float64, 3), "flags": (np.uint8, 1)}
        bufs = {}
        for k in want:
            bufs[k] = _Buf(out[k], *spec[k]) if out is not None and k in out else _alloc_like(q, q.n, *spec[k])
        p = lambda k: bufs[k].ptr if k in bufs else None   # noqa: E731
        self._check(self._lib.b200aabb_first_intersection(self._h, q.ptr, q.n, int(q.device), p("prim"), p("t"), p("point"), p("point2"), p("flags")))
        return {k: b.obj for k, b in bufs.items()}

    def first_intersection_skip(self, rays, skip_ids=None, skip_mask=None, want=("prim", "t", "point", "point2", "flags")):
        """first_intersection(Ray_3, skip): skip_ids = one primitive id per ray to ignore (NO_PRIMITIVE: none), skip_mask = bool array indexed
        by primitive id (True: ignore).  HOST batches."""
        q = _Buf(np.asarray(rays), np.float64, 6)
        spec = {"prim": (np.uint32, 1), "t": (np.float64, 1), "point": (np.float64, 3), "point2": (np.float64, 3), "flags": (np.uint8, 1)}
        bufs = {k: _alloc_like(q, q.n, *spec[k]) for k in want}
        p = lambda k: bufs[k].ptr if k in bufs else None   # noqa: E731
        ids = None if skip_ids is None else np.ascontiguousarray(skip_ids, dtype=np.uint32)
        words, nwords = None, 0
        if skip_mask is not None:
            bits = np.packbits(np.asarray(skip_mask, dtype=bool), bitorder="little")
            bits = np.concatenate([bits, np.zeros((-len(bits)) % 4, np.uint8)])
            words = np.ascontiguousarray(bits.view(np.uint32))
            nwords = len(words)
        self._check(self._lib.b200aabb_first_intersection_skip(self._h, q.ptr, q.n, 0, None if ids is None else ids.ctypes.data, None if words is None else words.ctypes.data,
                                                               nwords, p("prim"), p("t"), p("point"), p("point2"), p("flags")))
        return {k: b.obj for k, b in bufs.items()}

    def first_intersected_primitive(self, rays):
        return self.first_intersection(rays, want=("prim",))["prim"]

    # ---- distance queries (AABB_tree.h:363-377,464-477) ----
    def closest_point_and_primitive(self, points, want=("point", "prim", "sqdist", "flags"), out=None):
        q = _Buf(points, np.float64, 3)
        spec = {"point": (np.float64, 3), "prim": (np.uint32, 1), "sqdist": (np.float64, 1), "flags": (np.uint8, 1)}
        bufs = {}
        for k in want:
            bufs[k] = _Buf(out[k], *spec[k]) if out is not None and k in out else _alloc_like(q, q.n, *spec[k])
        p = lambda k: bufs[k].ptr if k in bufs else None   # noqa: E731
        self._check(self._lib.b200aabb_closest_point(self._h, q.ptr, q.n, int(q.device), p("point"), p("prim"), p("sqdist"), p("flags")))
        return {k: b.obj for k, b in bufs.items()}

    def closest_point(self, points):
        return self.closest_point_and_primitive(points, want=("point",))["point"]

    def squared_distance(self, points):
        return self.closest_point_and_primitive(points, want=("sqdist",))["sqdist"]

    def max_distance_to_mesh(self, points):
        """PMP::max_distance_to_mesh's inner loop (distance.h:98-190): (largest distance from the sample points to the mesh, index of a farthest sample)"""
        q = _Buf(points, np.float64, 3)
        d, i = C.c_double(), C.c_uint64()
        self._check(self._lib.b200aabb_max_distance_to_mesh(self._h, q.ptr, q.n, int(q.device), C.byref(d), C.byref(i)))
        return d.value, i.value

    def side_of_mesh(self, points, want_attempts=False, xf=None):
        """Side_of_triangle_mesh::operator() for a CLOSED mesh (Side_of_triangle_mesh.h:265-306): int8 per point, +1 ON_BOUNDED_SIDE, 0 ON_BOUNDARY,
        -1 ON_UNBOUNDED_SIDE; with want_attempts also the number of rays cast per point.  xf: the 3 x 4 affine map the mesh stands under (the points are in the
        frame of the transformed mesh), as in Rigid_triangle_mesh_collision_detection::does_A_contains_a_CC_of_B"""
        q = _Buf(points, np.float64, 3)
        o = _alloc_like(q, q.n, np.int8, 1)
        a = _alloc_like(q, q.n, np.uint8, 1) if want_attempts else None
        if xf is not None:
            m = np.ascontiguousarray(xf, np.float64).reshape(12)
            self._check(self._lib.b200aabb_side_of_mesh_transformed(self._h, m.ctypes.data, q.ptr, q.n, int(q.device), o.ptr, a.ptr if a is not None else None))
            return (o.obj, a.obj) if want_attempts else o.obj
        self._check(self._lib.b200aabb_side_of_mesh(self._h, q.ptr, q.n, int(q.device), o.ptr, a.ptr if a is not None else None))
        return (o.obj, a.obj) if want_attempts else o.obj

    # ---- multi-GPU plumbing: the built tree as flat device buffers (for torch.distributed.broadcast) ----
    def export_layout(self):
        ptrs = (C.c_void_p * 8)()
        sizes = (C.c_uint64 * 8)()
        cnt = C.c_int()
        self._check(self._lib.b200aabb_export_layout(self._h, ptrs, sizes, 8, C.byref(cnt)))
        return [(ptrs[i] or 0, sizes[i]) for i in range(cnt.value)]

    def import_alloc(self, n):
        self._check(self._lib.b200aabb_import_alloc(self._h, int(n)))

    def import_finish(self):
        self._check(self._lib.b200aabb_import_finish(self._h))


class Comm:
    """b200aabb_comm: one NCCL communicator rank (include/b200aabb.h, multi-GPU section).  One process per GPU: rank 0 makes
    Comm.unique_id(), hands the 128 bytes to the other ranks, every rank constructs Comm(device, rank, world, uid).  One process driving
    several GPUs: Comm.init_all(devices) and broadcast_tree_all."""

    def __init__(self, device, rank, world, uid, _handle=None):
        self._lib = load_library()
        self.device, self.rank, self.world = int(device), int(rank), int(world)
        if _handle is not None:
            self._h = _handle
            return
        h = C.c_void_p()
        rc = self._lib.b200aabb_comm_init(self.device, self.rank, self.world, C.c_char_p(bytes(uid)), C.byref(h))
        if rc:
            raise B200AABBError(rc)
        self._h = h

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        rc = load_library().b200aabb_comm_unique_id(buf)
        if rc:
            raise B200AABBError(rc)
        return buf.raw

    @classmethod
    def init_all(cls, devices):
        L = load_library()
        n = len(devices)
        devs = (C.c_int * n)(*[int(d) for d in devices])
        hs = (C.c_void_p * n)()
        rc = L.b200aabb_comm_init_all(n, devs, hs)
        if rc:
            raise B200AABBError(rc)
        return [cls(devices[i], i, n, None, _handle=C.c_void_p(hs[i])) for i in range(n)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b200aabb_comm_destroy(self._h)
            self._h = None

    def broadcast_tree(self, tree, root=0, mode=0):
        """replicates rank `root`'s tree into `tree` on every other rank; returns (device ms, bytes moved) of this rank"""
        ms, nb = C.c_double(), C.c_uint64()
        rc = self._lib.b200aabb_comm_broadcast_tree(self._h, tree._h, int(root), int(mode), C.byref(ms), C.byref(nb))
        if rc:
            raise B200AABBError(rc, (self._lib.b200aabb_comm_last_error(self._h) or b"").decode() or self._lib.b200aabb_last_error(tree._h).decode())
        return ms.value, nb.value


def broadcast_tree_all(comms, trees, root=0, mode=0):
    """single-process form: comms = Comm.init_all(devices), trees[i] lives on devices[i]; returns (ms of the slowest rank, bytes)"""
    L = load_library()
    n = len(comms)
    cs = (C.c_void_p * n)(*[c._h for c in comms])
    ts = (C.c_void_p * n)(*[t._h for t in trees])
    ms, nb = C.c_double(), C.c_uint64()
    rc = L.b200aabb_comm_broadcast_tree_all(cs, ts, n, int(root), int(mode), C.byref(ms), C.byref(nb))
    if rc:
        raise B200AABBError(rc, (L.b200aabb_comm_last_error(comms[0]._h) or b"").decode())
    return ms.value, nb.value
