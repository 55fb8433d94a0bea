"""diagnostic (GPU box): how far apart are the library's and the oracle's answers when their ids differ?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cgal_b200 as B, orc, bench
from cgal_b200 import workloads as W

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
v, f = W.terrain(nx, ny)
g, o = B.AABB_tree(v, f), orc.OracleTree(v, f)
rays = np.ascontiguousarray(bench.shard_rays(0)[::2]) if nx == 4096 else W.pinhole_rays(2048, 2048, nx, ny)
gr, rr = g.first_intersection(rays), o.first_intersection(rays)
hit = rr["prim"] != B.NO_PRIMITIVE
assert np.array_equal(hit, gr["prim"] != B.NO_PRIMITIVE)
mis = gr["prim"] != rr["prim"]
rel = np.abs(gr["t"] - rr["t"]) / np.maximum(np.abs(rr["t"]), 1e-300)
ulp = np.abs(gr["t"] - rr["t"]) / np.spacing(np.abs(rr["t"]))
print("rays", len(rays), "hit", hit.sum(), "id mismatches", mis.sum(), "t differs", (gr["t"] != rr["t"]).sum(), "gpu t > ref t", (gr["t"] > rr["t"]).sum())
print("flags among mismatches: tie", ((gr["flags"][mis] & 4) != 0).sum(), "near", ((gr["flags"][mis] & 16) != 0).sum())
print("flagged overall: tie", ((gr["flags"] & 4) != 0).sum(), "near", ((gr["flags"] & 16) != 0).sum())
print("ulp gap of mismatches, sorted:", np.sort(ulp[mis])[-40:])
print("rel gap max", rel.max())
pts = W.bbox_points(g.bbox(), 1 << 19)
gc, rc = g.closest_point_and_primitive(pts), o.closest_point(pts)
mis = gc["prim"] != rc["prim"]
print("points", len(pts), "id mismatches", mis.sum(), "sqd differs", (gc["sqdist"] != rc["sqd"]).sum())
print("flags among mismatches: tie", ((gc["flags"][mis] & 4) != 0).sum(), "near", ((gc["flags"][mis] & 16) != 0).sum())
print("flagged overall: tie", ((gc["flags"] & 4) != 0).sum(), "near", ((gc["flags"] & 16) != 0).sum())
u = np.abs(gc["sqdist"] - rc["sqd"]) / np.spacing(rc["sqd"])
print("ulp gap of sqd where it differs:", np.sort(u[u > 0])[-40:])

# ---- anatomy of the ray mismatches
mis = np.nonzero(gr["prim"] != rr["prim"])[0]
sub = rays[mis]
naive = o.naive_first_intersection(sub)
print("naive (all primitives, min rounded t, lowest id): agrees with gpu", (naive["prim"] == gr["prim"][mis]).sum(), "with ref", (naive["prim"] == rr["prim"][mis]).sum(), "ntie", naive["ntie"][:20])
np.set_printoptions(precision=17, linewidth=250)
for k in range(min(12, len(mis))):
    i = mis[k]
    kg, tg, _ = o.candidate_ray_hit(gr["prim"][i], rays[i])
    kr, tr, _ = o.candidate_ray_hit(rr["prim"][i], rays[i])
    print("ray", rays[i], "gpu prim", gr["prim"][i], "t", gr["t"][i], "flags", gr["flags"][i], "| ref prim", rr["prim"][i], "t", rr["t"][i],
          "| oracle per-triangle: gpu's tri hit", kg, tg, "ref's tri hit", kr, tr, "| naive", naive["prim"][k], naive["t"][k])
    print("   gpu tri", v[f[gr["prim"][i]]].ravel(), "ref tri", v[f[rr["prim"][i]]].ravel())

# ---- device predicates on the failing (triangle, ray) pairs
from cgal_b200.aabb_tree import debug_predicate
recs = np.concatenate([v[f[gr["prim"][mis]].astype(np.int64)].reshape(-1, 9), rays[mis]], 1)
res, ex, val = debug_predicate(6, recs)
print("device Intersect_3 on (gpu's tri, ray): kind", res, "used_exact", ex)
for k in range(min(6, len(mis))):
    tri, ray = recs[k, :9].reshape(3, 3), recs[k, 9:]
    s_, q_ = ray[:3], ray[3:]
    a, b, c = tri
    d_ = q_ - s_
    shifted = a + d_
    quads = {"abc,shifted": (a, b, c, shifted), "abc,s": (a, b, c, s_), "s,q,a,b": (s_, q_, a, b), "s,q,b,c": (s_, q_, b, c), "s,q,c,a": (s_, q_, c, a),
             "q,s,a,b": (q_, s_, a, b), "q,s,b,c": (q_, s_, b, c), "q,s,c,a": (q_, s_, c, a)}
    for name, quad in quads.items():
        rec = np.concatenate(quad)[None, :]
        dv, dex, _ = debug_predicate(0, rec)
        de, _, _ = debug_predicate(8, rec)
        print(k, name, "device filtered", dv[0], "used_exact", dex[0], "device exact", de[0], "oracle", orc.orientation(*quad), "oracle exact", orc.orientation(*quad, exact=True))
