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
    # both meshes under transformations (Rigid_triangle_mesh_collision_detection): the same tree twice, the copy turned by 0.3 rad about z and scaled by 1.0001
    # (nested, nothing intersects) / moved by half a radius (intersects); inclusion test of 1 Mi points against the turned mesh
    c, s_ = np.cos(0.3), np.sin(0.3)
    turn = np.array([[c, -s_, 0, 0], [s_, c, 0, 0], [0, 0, 1, 0]], float)
    xt = {}
    for name, m in (("turned_and_scaled_1.0001", np.hstack([turn[:, :3] * 1.0001, np.zeros((3, 1))])), ("turned_and_shifted", np.hstack([turn[:, :3], np.array([[0.5], [0.0], [0.0]])]))):
        ms, r = timed(lambda: g.do_intersect_tree(g, other_xf=m))
        xt[name] = {"ms": ms, "result": bool(r)}
    pts1m = torch.from_numpy(rng.uniform(-1.1, 1.1, (1 << 20, 3))).cuda()
    ms, side = timed(lambda: g.side_of_mesh(pts1m, xf=turn))
    ms0, side0 = timed(lambda: g.side_of_mesh(pts1m))
    xt["side_of_mesh_1Mi_points"] = {"ms_under_a_rotation": ms, "ms_plain": ms0, "inside_fraction": float((side == 1).float().mean()), "inside_fraction_plain": float((side0 == 1).float().mean())}
    out["under_transformations_icosphere8"] = xt
    print(json.dumps(xt), flush=True)
    # Segment_3 primitives: the 1,966,080 edges of the sphere; closest_point of 1 Mi points (one thread per query), Plane_3 counts
    e = np.unique(np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64), axis=1), axis=0)
    gs = B.AABB_tree(segments=np.concatenate([v[e[:, 0]], v[e[:, 1]]], axis=1))
    gs.build()
    ms, r = timed(lambda: gs.closest_point_and_primitive(pts1m))
    nrm = rng.normal(size=(1 << 16, 3))
    planes = torch.from_numpy(np.concatenate([nrm, rng.uniform(-1, 1, (1 << 16, 1)) * np.linalg.norm(nrm, axis=1, keepdims=True)], axis=1)).cuda()
    msp, cnt = timed(lambda: gs.number_of_intersected_primitives(planes, B.PLANE))
    o_s = orc.OracleTree(segments=np.concatenate([v[e[:, 0]], v[e[:, 1]]], axis=1))
    smp = pts1m[:20000].cpu().numpy()
    want = o_s.closest_point(smp)
    out["segment_primitives_icosphere8_edges"] = {"segments": len(e), "closest_point_ms_1Mi": ms, "Mqueries_s": (1 << 20) / ms / 1e3, "plane_count_ms_64Ki": msp,
                                                  "mean_edges_per_plane": float(cnt.float().mean()), "sample_sqdist_equal_to_oracle": bool(np.array_equal(r["sqdist"][:20000].cpu().numpy(), want["sqd"]))}
    print(json.dumps(out["segment_primitives_icosphere8_edges"]), flush=True)
    del g, gs
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
