"""Regenerates tests/golden/meshes.npz from the reference's fixture meshes (run in the build container only;
/root/reference does not exist on the GPU box).  Faces keep OFF file order = Surface_mesh face order
(SURVEY.md §8a row 5).  Usage: python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

REF = "/root/reference"
MESHES = {
    "elephant": "Data/data/meshes/elephant.off",            # BASELINE.json configs[0]
    "tetrahedron": "Data/data/meshes/tetrahedron.off",      # aabb_correctness_triangle_test.cpp:41
    "pinion_small": "Data/data/meshes/pinion_small.off",
    "handle": "Data/data/meshes/handle.off",
    "cube": "AABB_tree/test/AABB_tree/data/cube.off",
    "coverrear": "AABB_tree/test/AABB_tree/data/coverrear.off",
    "finger": "AABB_tree/test/AABB_tree/data/finger.off",
}


def read_off(path):
    toks = []
    with open(path) as f:
        for line in f:
            line = line.split("#")[0].strip()
            if line:
                toks.extend(line.split())
    assert toks[0] in ("OFF", "COFF", "NOFF"), toks[0]
    nv, nf = int(toks[1]), int(toks[2])
    pos = 4
    verts = np.array(toks[pos:pos + 3 * nv], dtype=np.float64).reshape(nv, 3)
    pos += 3 * nv
    faces = []
    for _ in range(nf):
        k = int(toks[pos])
        idx = [int(x) for x in toks[pos + 1:pos + 1 + k]]
        pos += 1 + k
        assert k == 3, "triangle meshes only"
        faces.append(idx)
    return verts, np.array(faces, dtype=np.uint32)


if __name__ == "__main__":
    out = {}
    for name, rel in MESHES.items():
        v, f = read_off(os.path.join(REF, rel))
        out[name + "_v"] = v
        out[name + "_f"] = f
        print(name, v.shape, f.shape, file=sys.stderr)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes.npz"), **out)
