"""The C++ host side (include/CGAL_B200/AABB_tree.h, the AABB_tree-compatible class template) compiles warning-free against
the C ABI and links to libb200aabb.so; on a GPU box (-m gpu) its known-answer tests run and its batched answers are
compared with the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import cpph


def test_cpp_class_compiles_and_links():
    exe = cpph.build()
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libb200aabb.so" in out and "not found" not in out.split("libb200aabb.so")[1].split("\n")[0]


def test_cpp_class_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([cpph.build()], capture_output=True, text=True)
    assert r.returncode == 2 and "b200aabb (-5)" in r.stderr     # CGAL_B200::Error(B200AABB_ERR_NO_DEVICE): no CPU fallback


@pytest.mark.gpu
def test_cpp_known_answers():
    r = subprocess.run([cpph.build()], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "all known-answer and consistency checks passed" in r.stdout


@pytest.mark.gpu
def test_cpp_batched_answers_match_oracle(meshes, tmp_path):
    import orc
    from cgal_b200 import workloads as W
    from cgal_b200 import FLAG_TIE, FLAG_NEAR_TIE  # noqa: F401
    v, f = meshes["elephant"]
    o = orc.OracleTree(v, f)
    bb = o.bbox()
    rays, segs, pts = W.sphere_rays(bb, 3000), W.bbox_segments(bb, 3000), W.bbox_points(bb, 3000)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as fh:
        fh.write(struct.pack("<5Q", len(v), len(f), len(rays), len(segs), len(pts)))
        for a in (np.asarray(v, np.float64), np.asarray(f, np.uint32), rays, segs, pts):
            fh.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([cpph.build(), fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(fout, "rb").read()
    ray_t = np.dtype([("id", "<u4"), ("p", "<f8", 3)])
    seg_t = np.dtype([("cnt", "<u4"), ("hit", "<u4")])
    pt_t = np.dtype([("id", "<u4"), ("cp", "<f8", 3), ("sd", "<f8")])
    off = 0
    gr = np.frombuffer(raw, ray_t, len(rays), off); off += ray_t.itemsize * len(rays)
    gs = np.frombuffer(raw, seg_t, len(segs), off); off += seg_t.itemsize * len(segs)
    gp = np.frombuffer(raw, pt_t, len(pts), off)
    ref_r, ref_p, ref_c = o.first_intersection(rays), o.closest_point(pts), o.count(segs, 0)
    np.testing.assert_array_equal(gs["cnt"], ref_c)
    np.testing.assert_array_equal(gs["hit"], (ref_c > 0).astype(np.uint32))
    same = gr["id"] == ref_r["prim"]
    assert same.mean() > 0.999                                        # the rest are reported ties (checked through the C ABI tests)
    np.testing.assert_array_equal(gr["p"][same], ref_r["pts"][same][:, :3])
    # ids differ from the reference's only on exact-distance ties (a closest point on a shared edge or vertex is the SAME point
    # for both triangles: the reference keeps the first it visited, this library the smallest id)
    same = gp["id"] == ref_p["prim"]
    np.testing.assert_array_equal(gp["cp"][same], ref_p["cp"][same])
    np.testing.assert_array_equal(gp["sd"][same], ref_p["sqd"][same])
    same_point = np.all(gp["cp"] == ref_p["cp"], axis=1)
    assert same_point.mean() > 0.999 and np.all(same | same_point | (gp["sd"] == ref_p["sqd"]))
    assert np.all(np.abs(gp["sd"] - ref_p["sqd"]) <= 1e-12 * ref_p["sqd"])
    for i in np.nonzero(~same)[0][:100]:        # the reported primitive really attains the reported point
        np.testing.assert_array_equal(o.candidate_closest(gp["id"][i], pts[i]), gp["cp"][i])
