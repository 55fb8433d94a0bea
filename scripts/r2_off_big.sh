#!/bin/bash
B200AABB_INGEST_TIMING=1 python - <<'PY'
# OFF ingest at scale: an 8.4 M-triangle terrain as "%.17g" text
import sys, time, io
sys.path.insert(0, '.')
import numpy as np
import cgal_b200 as B
from cgal_b200 import workloads as W
v, f = W.terrain(2048, 2048)
t0 = time.perf_counter()
buf = io.BytesIO()
buf.write(b"OFF\n%d %d 0\n" % (len(v), len(f)))
np.savetxt(buf, v, fmt="%.17g")
np.savetxt(buf, np.concatenate([np.full((len(f), 1), 3), f.astype(np.int64)], axis=1), fmt="%d")
text = buf.getvalue()
print("text of %d vertices / %d facets: %.1f MB (written in %.1f s)" % (len(v), len(f), len(text) / 1e6, time.perf_counter() - t0), flush=True)
g = B.AABB_tree()
for _ in range(3):
    t0 = time.perf_counter(); g.insert_off(text); t1 = time.perf_counter(); g.build(); t2 = time.perf_counter()
    print("device ingest call %.1f ms + build %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
h = B.AABB_tree(v, f)
print("size", g.size(), "bbox equal:", bool(np.array_equal(g.bbox(), h.bbox())))
bb = g.bbox()
rays, pts = W.sphere_rays(bb, 200000), W.bbox_points(bb, 200000)
a, b = g.first_intersection(rays), h.first_intersection(rays)
print("first_intersection equal:", all(bool(np.array_equal(a[k], b[k])) for k in ("prim", "t", "point", "flags")))
a, b = g.closest_point_and_primitive(pts), h.closest_point_and_primitive(pts)
print("closest_point equal:", all(bool(np.array_equal(a[k], b[k])) for k in ("prim", "sqdist", "point", "flags")))
PY
