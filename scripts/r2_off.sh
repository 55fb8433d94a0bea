#!/bin/bash
mkdir -p gpurun_out
B200AABB_INGEST_TIMING=1 python - <<'PY'
# timing at scale: a 1 M-triangle icosphere written as "%.17g" text (about 60 MB): device ingest against numpy's text reader + set_mesh
import sys, time, io
sys.path.insert(0, '.')
import numpy as np
import cgal_b200 as B
from cgal_b200 import workloads as W
v, f = W.icosphere(8)
t0 = time.perf_counter()
buf = io.StringIO()
buf.write("OFF\n%d %d 0\n" % (len(v), len(f)))
np.savetxt(buf, v, fmt="%.17g")
np.savetxt(buf, np.concatenate([np.full((len(f), 1), 3), f.astype(np.int64)], axis=1), fmt="%d")
text = buf.getvalue().encode()
print("text of %d vertices / %d facets: %.1f MB (written in %.1f s)" % (len(v), len(f), len(text) / 1e6, time.perf_counter() - t0))
g = B.AABB_tree()
for _ in range(5):
    t0 = time.perf_counter(); g.insert_off(text); t1 = time.perf_counter(); g.build(); t2 = time.perf_counter()
    print("device ingest %.1f ms (%.2f GB/s of text) + build %.2f ms" % ((t1 - t0) * 1e3, len(text) / (t1 - t0) / 1e9, (t2 - t1) * 1e3))
t0 = time.perf_counter()
arr = np.loadtxt(io.BytesIO(text), skiprows=2, max_rows=len(v))
print("numpy.loadtxt of the vertex block alone: %.2f s" % (time.perf_counter() - t0))
h = B.AABB_tree(arr, f)
print("bbox equal:", bool(np.array_equal(g.bbox(), h.bbox())), " vertices equal to the written doubles:", bool(np.array_equal(arr, v)))
pts = W.bbox_points(g.bbox(), 200000)
a, b = g.closest_point_and_primitive(pts), h.closest_point_and_primitive(pts)
print("closest_point results equal:", all(bool(np.array_equal(a[k], b[k])) for k in ("prim", "sqdist", "point", "flags")))
PY
