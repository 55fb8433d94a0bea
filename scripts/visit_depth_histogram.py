#!/usr/bin/env python
"""At which depths of the tree does a closest_point query fetch its nodes?  (CPU, the device traversal code compiled for the host.)
Answers how much a shared-memory copy of the top k levels could save: the share of fetches at depth < k.
  python scripts/visit_depth_histogram.py [grid_x grid_y queries]     # terrain of 2*grid_x*grid_y triangles (default 1024 x 512 = 1 Mi)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hc                                    # noqa: E402
from cgal_b200 import workloads as W         # noqa: E402

gx, gy, nq = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1024, 512, 20000)
v, f = W.terrain(gx, gy)
tree = hc.HostTree(v[f.astype(np.int64)].reshape(-1, 9))
bb = np.concatenate([v.min(0), v.max(0)])
pts = W.bbox_points(bb, nq)
hist = np.zeros(64, np.uint64)
depth = hc.lib().hc_closest_depth_histogram(tree.h, pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts), hist.ctypes.data_as(C.POINTER(C.c_uint64)))
per_query = hist / len(pts)
total = per_query.sum()
print("terrain %d triangles, %d bbox points, tree depth %d, %.1f node fetches per query" % (len(f), len(pts), depth, total))
print("depth  fetches/query  cumulative share")
acc = 0.0
for d in range(depth + 1):
    acc += per_query[d]
    print("%5d  %13.2f  %15.1f %%" % (d, per_query[d], 100 * acc / total))
