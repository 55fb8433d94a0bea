#!/usr/bin/env python
"""Small closest_point / intersection run for compute-sanitizer (memcheck, racecheck): elephant mesh, a few thousand queries."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    v, f = z["elephant_v"], z["elephant_f"]
    g = B.AABB_tree(v, f)
    bb = g.bbox()
    pts = W.bbox_points(bb, 6000, margin=0.3)
    r = g.closest_point_and_primitive(pts)
    rays = W.sphere_rays(bb, 4000)
    fi = g.first_intersection(rays)
    c = g.number_of_intersected_primitives(rays, B.RAY)
    off, ids = g.all_intersected_primitives(rays[:1000], B.RAY)
    print("ok", int(r["prim"].sum() % 1000), int((fi["prim"] != B.NO_PRIMITIVE).sum()), int(c.sum()), len(ids))


if __name__ == "__main__":
    main()
