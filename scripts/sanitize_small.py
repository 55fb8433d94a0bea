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
    # round 2: kernel 2, skip mask, segment / line / plane objects, side_of_mesh, max_distance, refit, a 20,000-triangle build (80 blocks of the fused kernel)
    g.set_option(B.OPT_CLOSEST_KERNEL, 2)
    r2 = g.closest_point_and_primitive(pts[:2000])
    g.set_option(B.OPT_CLOSEST_KERNEL, 1)
    assert np.array_equal(r2["sqdist"], r["sqdist"][:2000])
    sk = g.first_intersection_skip(rays[:1000], skip_ids=fi["prim"][:1000].copy())
    segs = W.bbox_segments(bb, 2000)
    o1 = g.all_intersections(segs, kind=B.SEGMENT)
    lines = np.concatenate([segs[:, :3], segs[:, 3:] - segs[:, :3]], 1)
    o2 = g.all_intersections(lines, kind=B.LINE)
    planes = np.concatenate([np.random.default_rng(1).normal(size=(200, 3)), np.zeros((200, 1))], 1)
    o3 = g.all_intersections(planes, kind=B.PLANE)
    side = g.side_of_mesh(pts)
    md = g.max_distance_to_mesh(pts)
    g.refit(vertices=v * 1.001, faces=f)
    soup = W.triangle_soup(20000)
    t = B.AABB_tree(triangles=soup)
    t.build()
    t.set_option(103, 1)
    t.build()
    # Triangle_3 queries and tree-vs-tree (plain and under transformations)
    tq = np.concatenate([pts[:3000], pts[:3000] + 0.05, pts[:3000] + np.array([0.05, -0.03, 0.02])], 1)
    ct = g.number_of_intersected_primitives(tq, B.TRIANGLE)
    offt, idst = g.all_intersected_primitives(tq, B.TRIANGLE)
    g2 = B.AABB_tree(v + np.array([0.1, 0.0, 0.0]), f)
    tt = g.do_intersect_tree(g2), g.do_intersect_tree(g2, xf=np.array([[0, -1, 0, 0.2], [1, 0, 0, 0], [0, 0, 1, 0]], float)), g.do_intersect_tree(g2, other_xf=np.array([[1, 0, 0, 9.0], [0, 1, 0, 0], [0, 0, 1, 0]], float))
    print("triangle queries", int(ct.sum()), len(idst), tt)
    # Segment_3 primitives: the edges of the mesh
    e = np.unique(np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64), axis=1), axis=0)
    gs = B.AABB_tree(segments=np.concatenate([v[e[:, 0]], v[e[:, 1]]], axis=1))
    rs = gs.closest_point_and_primitive(pts[:3000])
    cs = gs.number_of_intersected_primitives(tq, B.TRIANGLE), gs.number_of_intersected_primitives(planes, B.PLANE)
    print("segment primitives", gs.size(), float(rs["sqdist"].sum()), int(cs[0].sum()), int(cs[1].sum()))
    print("ok", int(r["prim"].sum() % 1000), int((fi["prim"] != B.NO_PRIMITIVE).sum()), int(c.sum()), len(ids), int((sk["prim"] != B.NO_PRIMITIVE).sum()), len(o1[1]), len(o2[1]),
          len(o3[1]), int((side == 1).sum()), round(md[0], 6), t.size())


if __name__ == "__main__":
    main()
