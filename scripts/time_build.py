#!/usr/bin/env python
"""GPU-side build timing (terrain16M and a 32M soup): device time of b200aabb_build, best of 5.  Not a bench value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.build()
    ts = []
    for _ in range(5):
        g.build()
        ts.append(g.stats()["last_build_ms"])
    print("terrain16M build ms:", ["%.3f" % t for t in ts], "-> %.0f Mtri/s" % (len(f) / min(ts) / 1e3), flush=True)
    import time
    wall = []
    for _ in range(5):
        t0 = time.perf_counter(); g.build(); wall.append((time.perf_counter() - t0) * 1e3)
    print("terrain16M build, host wall clock of b200aabb_build (allocations + kernels + final synchronize) ms:", ["%.3f" % t for t in wall], flush=True)
    # refit-only rebuild (b200aabb_refit_mesh: same topology, moved vertices) and the consumers of the query path, on the same tree
    import numpy as np
    import torch
    v2 = v.copy()
    v2[:, 2] *= 1.01
    ts = []
    for _ in range(4):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record(); g.refit(vertices=v2, faces=f); t1.record(); torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1))
    print("terrain16M refit_mesh ms (host vertices, incl. the 403 MB upload):", ["%.2f" % t for t in ts], flush=True)
    soup_d = torch.from_numpy(np.ascontiguousarray(v2[f.astype(np.int64)].reshape(-1, 9))).cuda()
    ts = []
    for _ in range(4):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record(); g.refit(triangles=soup_d); t1.record(); torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1))
    print("terrain16M refit ms (device triangles):", ["%.2f" % t for t in ts], flush=True)
    pts = torch.from_numpy(W.bbox_points(g.bbox(), 1 << 22)).cuda()
    for name, fn in (("max_distance_to_mesh", lambda: g.max_distance_to_mesh(pts)), ("side_of_mesh (open terrain: timing only)", lambda: g.side_of_mesh(pts))):
        fn()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record(); fn(); t1.record(); torch.cuda.synchronize()
        print("terrain16M %s, 4 Mi points: %.2f ms = %.0f Mq/s" % (name, t0.elapsed_time(t1), (1 << 22) / t0.elapsed_time(t1) / 1e3), flush=True)
    del soup_d, pts
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
    s = B.AABB_tree(triangles=W.triangle_soup(n))
    s.build()
    ts = []
    for _ in range(3):
        s.build()
        ts.append(s.stats()["last_build_ms"])
    print("soup %d build ms:" % n, ["%.3f" % t for t in ts], "-> %.0f Mtri/s" % (n / min(ts) / 1e3), flush=True)


if __name__ == "__main__":
    main()
