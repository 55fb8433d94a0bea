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


def check_first_intersection(gpu, ref, oracle_tree, rays, max_mismatch_frac=1e-4):
    """hit/no-hit equal; same id -> t and point bit-equal; id mismatches must be REPORTED (exact tie, or near tie = inside the
    rounding uncertainty of the winner, where the reference's own answer is an artefact of its traversal order), justified
    (the reference's primitive really is hit, at a t within 1e-5 relative) and rare."""
    from cgal_b200 import FLAG_HIT, FLAG_NEAR_TIE, FLAG_TIE, NO_PRIMITIVE
    hit_g = gpu["prim"] != NO_PRIMITIVE
    hit_r = ref["prim"] != NO_PRIMITIVE
    np.testing.assert_array_equal(hit_g, hit_r)
    np.testing.assert_array_equal((gpu["flags"] & FLAG_HIT) != 0, hit_g)
    same = gpu["prim"] == ref["prim"]
    tie = (gpu["flags"] & FLAG_TIE) != 0
    near = (gpu["flags"] & FLAG_NEAR_TIE) != 0
    np.testing.assert_array_equal(gpu["t"][same], ref["t"][same])                 # bit-exact ray parameter
    np.testing.assert_array_equal(gpu["point"][same], ref["pts"][same][:, :3])    # bit-exact hit point
    assert np.all(same | tie | near), "id mismatch without a reported tie: %d" % int((~(same | tie | near)).sum())
    # an id that differs on an EXACT tie carries the same t, bit for bit; only rows flagged NEAR_TIE (the runner-up lies inside the rounding
    # uncertainty of the winner, where the reference's answer depends on its own traversal order) may differ in t, by at most that band
    exact_rows = ~same & ~near
    np.testing.assert_array_equal(gpu["t"][exact_rows], ref["t"][exact_rows])
    assert (~same & near & ~tie).sum() <= max(2, max_mismatch_frac * len(same)), "too many near-tie mismatches"
    for i in np.nonzero(~same)[0][:300]:                                          # the reference id must be in the (near-)tie set
        k, t, _ = oracle_tree.candidate_ray_hit(ref["prim"][i], rays[i])
        assert k and t == ref["t"][i] and (abs(t - gpu["t"][i]) <= 1e-5 * abs(t) if near[i] else t == gpu["t"][i])
    return int((~same).sum()), int((tie | near).sum())


def check_closest_point(gpu, ref, oracle_tree, pts):
    import orc
    from cgal_b200 import FLAG_NEAR_TIE, FLAG_TIE
    same = gpu["prim"] == ref["prim"]
    tie = (gpu["flags"] & FLAG_TIE) != 0
    near = (gpu["flags"] & FLAG_NEAR_TIE) != 0
    assert np.all(same | tie | near), "id mismatch without a reported tie"
    np.testing.assert_array_equal(gpu["point"][same], ref["cp"][same])  # bit-exact closest point
    np.testing.assert_array_equal(gpu["sqdist"][same], ref["sqd"][same])
    rel = np.abs(gpu["sqdist"] - ref["sqd"]) <= 1e-12 * np.maximum(ref["sqd"], 1e-300)
    assert np.all(rel), "squared distance off by more than 1e-12 relative"
    for i in np.nonzero(~same)[0][:200]:
        c = orc.compare_distance(pts[i], gpu["point"][i], ref["cp"][i])
        assert c == 0 or (near[i] and c < 0)                                        # exact-distance tie, or the true minimum inside the band
        cand = oracle_tree.candidate_closest(gpu["prim"][i], pts[i])          # the primitive's projection, or its reference point (the hint)
        assert np.array_equal(cand, gpu["point"][i]) or np.array_equal(oracle_tree.reference_point(gpu["prim"][i]), gpu["point"][i])
    return int((~same).sum()), int((tie | near).sum())
