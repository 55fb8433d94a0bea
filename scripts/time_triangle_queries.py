#!/usr/bin/env python
"""Triangle_3 queries and tree-vs-tree do_intersect at scale on one GPU (timing + checks that do not need the oracle at full size):
  * icosphere level 8 (1,310,720 triangles): 4 Mi random small Triangle_3 queries, do_intersect / count, device-resident, best of 3, with a
    100 k-query sample checked against the CPU oracle;
  * tree-vs-tree: the sphere against a copy scaled by 1.0001 (nested shells, no pair intersects: the worst case, every primitive is traversed),
    against a copy shifted by half a radius (intersecting: early exit) and against a far copy (root boxes disjoint); the terrain16M tree against
    the sphere placed across it.
  python scripts/time_triangle_queries.py [--out gpurun_out/r2_21_triangle_queries.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, reps=3):
    import torch
    best, out = None, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    return best, out


def main():
    import torch
    import cgal_b200 as B
    import orc
    from cgal_b200 import workloads as W
    out = {}
    v, f = W.icosphere(8)
    g = B.AABB_tree(v, f)
    g.build()
    rng = np.random.default_rng(7)
    n = 1 << 22
    ctr = rng.uniform(-1.2, 1.2, (n, 3))
    q = np.concatenate([ctr, ctr + rng.uniform(-0.02, 0.02, (n, 3)), ctr + rng.uniform(-0.02, 0.02, (n, 3))], axis=1)
    qd = torch.from_numpy(q).cuda()
    ms_any, hit = timed(lambda: g.do_intersect(qd, B.TRIANGLE))
    ms_cnt, cnt = timed(lambda: g.number_of_intersected_primitives(qd, B.TRIANGLE))
    st = g.stats()
    cnt = cnt.cpu().numpy()
    m = 100_000
    o = orc.OracleTree(v, f)
    t0 = time.perf_counter()
    want = o.count(q[:m], B.TRIANGLE)
    cpu_s = time.perf_counter() - t0
    out["triangle_queries_icosphere8"] = {"triangles": len(f), "queries": n, "do_intersect_ms": ms_any, "count_ms": ms_cnt, "Mqueries_s_do_intersect": n / ms_any / 1e3,
                                          "Mqueries_s_count": n / ms_cnt / 1e3, "hit_fraction": float((cnt > 0).mean()), "max_count": int(cnt.max()),
                                          "nodes_per_query": st["nodes_visited"] / n, "leaves_per_query": st["leaves_tested"] / n,
                                          "oracle_sample": m, "sample_counts_equal": bool(np.array_equal(cnt[:m], want)),
                                          "sample_do_intersect_equal": bool(np.array_equal(hit.cpu().numpy()[:m] != 0, want > 0)),
                                          "cpu_oracle_Mqueries_s": m / cpu_s / 1e6, "cpu_threads": orc.nthreads()}
    print(json.dumps(out["triangle_queries_icosphere8"]), flush=True)
    tt = {}
    for name, w in (("nested_shell_1.0001", v * 1.0001), ("shifted_half_radius", v + np.array([0.5, 0.0, 0.0])), ("far_copy", v + np.array([10.0, 0.0, 0.0]))):
        g2 = B.AABB_tree(w, f)
        g2.build()
        ms, r = timed(lambda: g.do_intersect_tree(g2, witness=True))
        tt[name] = {"ms": ms, "result": bool(r[0]), "witness": [int(r[1]), int(r[2])]}
        if r[0]:
            sa, sb = v[f[r[1]]].ravel(), w[f[r[2]]].ravel()
            tt[name]["witness_pair_intersects_per_oracle"] = bool(orc.tri_tri_do_intersect(sa, sb))
        del g2
    # the same nested shells decided by the oracle on a smaller sphere (level 6, 81,920 triangles each): the GPU answer must agree
    v6, f6 = W.icosphere(6)
    a6, b6 = B.AABB_tree(v6, f6), B.AABB_tree(v6 * 1.0001, f6)
    o6, p6 = orc.OracleTree(v6, f6), orc.OracleTree(v6 * 1.0001, f6)
    t0 = time.perf_counter()
    want6 = o6.do_intersect_tree(p6)
    tt["level6_nested_oracle_agrees"] = bool(a6.do_intersect_tree(b6) == want6)
    tt["level6_nested_oracle_seconds"] = time.perf_counter() - t0
    tt["level6_nested_result"] = bool(want6)
    out["tree_vs_tree_icosphere8"] = tt
    print(json.dumps(tt), flush=True)
    del g
    tv, tf = W.terrain()
    gt = B.AABB_tree(tv, tf)
    gt.build()
    zc = float(tv[:, 2].mean())
    big = {}
    for name, w in (("sphere_across_the_terrain", v * 200.0 + np.array([2048.0, 1024.0, zc])), ("sphere_high_above", v * 200.0 + np.array([2048.0, 1024.0, zc + 1000.0]))):
        g2 = B.AABB_tree(w, f)
        g2.build()
        ms, r = timed(lambda: gt.do_intersect_tree(g2))
        ms2, r2 = timed(lambda: g2.do_intersect_tree(gt))
        big[name] = {"ms": ms, "ms_swapped_arguments": ms2, "result": bool(r), "symmetric": bool(r == r2)}
        del g2
    out["tree_vs_tree_terrain16M_vs_icosphere8"] = big
    print(json.dumps(big), flush=True)
    if "--out" in sys.argv:
        with open(sys.argv[sys.argv.index("--out") + 1], "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
