"""helpers for the -m gpu tests"""
import ctypes as C

import numpy as np

_RT = None


def cudart():
    global _RT
    if _RT is None:
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                _RT = C.CDLL(name)
                break
            except OSError:
                continue
        if _RT is None:
            import torch  # noqa: F401  (loads a cudart)
            _RT = C.CDLL("libcudart.so.12")
        _RT.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    return _RT


def device_to_host(ptr, nbytes):
    out = np.zeros(nbytes, np.uint8)
    if nbytes:
        rc = cudart().cudaMemcpy(out.ctypes.data, C.c_void_p(ptr), nbytes, 2)
        assert rc == 0, "cudaMemcpy D2H failed: %d" % rc
    return out


def check_first_intersection(gpu, ref, oracle_tree, rays):
    """prim / t / point equal to the oracle except exact-t ties, which must be flagged and justified."""
    from cgal_b200 import FLAG_HIT, FLAG_TIE, NO_PRIMITIVE
    hit_g = gpu["prim"] != NO_PRIMITIVE
    hit_r = ref["prim"] != NO_PRIMITIVE
    np.testing.assert_array_equal(hit_g, hit_r)
    np.testing.assert_array_equal((gpu["flags"] & FLAG_HIT) != 0, hit_g)
    np.testing.assert_array_equal(gpu["t"], ref["t"])                 # bit-exact ray parameter
    same = gpu["prim"] == ref["prim"]
    tie = (gpu["flags"] & FLAG_TIE) != 0
    assert np.all(same | tie), "id mismatch without a reported tie"
    np.testing.assert_array_equal(gpu["point"][same], ref["pts"][same][:, :3])
    for i in np.nonzero(~same)[0][:200]:                              # the reference id must be in the tie set
        k, t, _ = oracle_tree.candidate_ray_hit(ref["prim"][i], rays[i])
        assert k and t == gpu["t"][i]
    return int((~same).sum()), int(tie.sum())


def check_closest_point(gpu, ref, oracle_tree, pts):
    import orc
    from cgal_b200 import FLAG_TIE
    same = gpu["prim"] == ref["prim"]
    tie = (gpu["flags"] & FLAG_TIE) != 0
    assert np.all(same | tie), "id mismatch without a reported tie"
    np.testing.assert_array_equal(gpu["point"][same], ref["cp"][same])  # bit-exact closest point
    np.testing.assert_array_equal(gpu["sqdist"][same], ref["sqd"][same])
    rel = np.abs(gpu["sqdist"] - ref["sqd"]) <= 1e-12 * np.maximum(ref["sqd"], 1e-300)
    assert np.all(rel), "squared distance off by more than 1e-12 relative"
    for i in np.nonzero(~same)[0][:200]:
        assert orc.compare_distance(pts[i], gpu["point"][i], ref["cp"][i]) == 0     # exact-distance tie
        np.testing.assert_array_equal(oracle_tree.candidate_closest(gpu["prim"][i], pts[i]), gpu["point"][i])
    return int((~same).sum()), int(tie.sum())
