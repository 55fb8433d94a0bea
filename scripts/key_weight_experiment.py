#!/usr/bin/env python
"""Tree quality experiment (CPU): node fetches of closest_point / first_intersection queries on a terrain as a function of the per-axis
weights of the Morton frame (HC_KEY_WEIGHTS=wx,wy,wz read by tests/hostcheck).  python scripts/key_weight_experiment.py [gx gy nq]"""
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
soup = v[f.astype(np.int64)].reshape(-1, 9)
bb = np.concatenate([v.min(0), v.max(0)])
pts = W.bbox_points(bb, nq)
for w in ("1,1,1", "1,1,0.5", "1,1,0.25", "1,1,0.0625", "1,1,0.00390625", "1,1,0"):
    os.environ["HC_KEY_WEIGHTS"] = w
    tree = hc.HostTree(soup)
    hist = np.zeros(64, np.uint64)
    depth = hc.lib().hc_closest_depth_histogram(tree.h, pts.ctypes.data_as(C.POINTER(C.c_double)), len(pts), hist.ctypes.data_as(C.POINTER(C.c_uint64)))
    print("weights %-16s depth %2d  closest_point node fetches per query %.1f" % (w, depth, hist.sum() / len(pts)), flush=True)
