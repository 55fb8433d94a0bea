"""cgal_b200 — B200-native AABB tree, drop-in for the hot path of CGAL::AABB_tree.

The product is libb200aabb.so (hand-written CUDA for sm_100a behind the C ABI of include/b200aabb.h).  This package
is the host-side mirror of the reference's AABB_tree interface used by the tests and the benchmark; it calls the C ABI
through ctypes and has NO CPU fallback: importing works without a GPU, any compute call without one raises.
"""
from .aabb_tree import AABB_tree, B200AABBError, SEGMENT, RAY, LINE, PLANE, TRIANGLE, FLAG_HIT, FLAG_SEGMENT, FLAG_TIE, FLAG_EXACT, FLAG_NEAR_TIE, FLAG_TRIANGLE, OPT_SORT_QUERIES, OPT_CLOSEST_KERNEL, OPT_WARP_EXACT_GO, OPT_HOST_PARTS, OPT_ASYNC_HOST, NO_PRIMITIVE, load_library, Comm, broadcast_tree_all  # noqa: F401

__all__ = ["AABB_tree", "B200AABBError", "SEGMENT", "RAY", "LINE", "PLANE", "TRIANGLE", "FLAG_HIT", "FLAG_SEGMENT", "FLAG_TIE", "FLAG_EXACT", "FLAG_NEAR_TIE", "FLAG_TRIANGLE", "OPT_SORT_QUERIES", "OPT_CLOSEST_KERNEL", "OPT_WARP_EXACT_GO", "OPT_HOST_PARTS", "OPT_ASYNC_HOST", "NO_PRIMITIVE", "load_library", "Comm", "broadcast_tree_all"]
