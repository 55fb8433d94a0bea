#!/usr/bin/env python
"""BASELINE.json configs 1-4 at FULL size on one GPU: throughput (device-resident queries, best of 3) and parity with the CPU oracle on the
subsets BASELINE.md §3 names (config 1: all 1 M + 1 M queries; config 2: a strided 1 M subset + three z-slices of the 256^3 grid;
config 3: 1 M primary rays + 1 M shadow segments; config 4: the 10 M-triangle prefix).  One JSON object per config on stdout and, with
--out, all of them in one file (committed under profiles/).  Not the bench line (bench.py measures configs[2]/[4]).

  python scripts/run_configs.py [1 2 3 4] [--out gpurun_out/r2_parity_configs.json] [--no-parity]
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
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


def ray_parity(B, g, o, rays_np):
    """first_intersection of the GPU (numpy in, HOST path) against the oracle: counts of every kind of disagreement"""
    got, ref = g.first_intersection(rays_np), o.first_intersection(rays_np)
    hit_g, hit_r = got["prim"] != B.NO_PRIMITIVE, ref["prim"] != B.NO_PRIMITIVE
    same = got["prim"] == ref["prim"]
    tie, near = (got["flags"] & B.FLAG_TIE) != 0, (got["flags"] & B.FLAG_NEAR_TIE) != 0
    rel = np.abs(got["t"] - ref["t"]) / np.maximum(np.abs(ref["t"]), 1e-300)
    return {"rays": len(rays_np), "hits": int(hit_r.sum()), "hit_flag_mismatch": int((hit_g != hit_r).sum()), "id_mismatch": int((~same).sum()),
            "id_mismatch_not_flagged": int((~same & ~tie & ~near).sum()), "flag_tie": int(tie.sum()), "flag_near_tie": int(near.sum()),
            "id_mismatch_on_near_tie": int((~same & near).sum()), "t_bit_equal_where_id_equal": bool(np.array_equal(got["t"][same], ref["t"][same])),
            "point_bit_equal_where_id_equal": bool(np.array_equal(got["point"][same], ref["pts"][same][:, :3])),
            "t_bit_equal_outside_near_ties": bool(np.array_equal(got["t"][~near], ref["t"][~near])), "max_rel_t_diff": float(rel.max()) if len(rel) else 0.0,
            "max_rel_t_diff_outside_near_ties": float(rel[~near].max()) if (~near).any() else 0.0}


def point_parity(B, g, o, pts_np):
    got, ref = g.closest_point_and_primitive(pts_np), o.closest_point(pts_np)
    same = got["prim"] == ref["prim"]
    tie = (got["flags"] & B.FLAG_TIE) != 0
    rel = np.abs(got["sqdist"] - ref["sqd"]) / np.maximum(ref["sqd"], 1e-300)
    return {"points": len(pts_np), "id_mismatch": int((~same).sum()), "id_mismatch_not_flagged": int((~same & ~tie).sum()), "flag_tie": int(tie.sum()),
            "sqdist_bit_equal": bool(np.array_equal(got["sqdist"], ref["sqd"])), "closest_point_bit_equal_where_id_equal": bool(np.array_equal(got["point"][same], ref["cp"][same])),
            "max_rel_sqdist_diff": float(rel.max()) if len(rel) else 0.0}


def main():
    import torch
    import cgal_b200 as B
    import orc
    from cgal_b200 import workloads as W
    argv = sys.argv[1:]
    out_path = None
    if "--out" in argv:
        out_path = argv[argv.index("--out") + 1]
        argv = [a for i, a in enumerate(argv) if a != "--out" and (i == 0 or argv[i - 1] != "--out")]
    parity = "--no-parity" not in argv
    which = [a for a in argv if a in ("1", "2", "3", "4")] or ["1", "2", "3", "4"]
    results = {}

    def emit(key, obj):
        results[key] = obj
        print(json.dumps({key: obj}), flush=True)

    def stream(g):
        g.set_stream(torch.cuda.current_stream().cuda_stream)

    if "1" in which:      # elephant, 1 M sphere rays + 1 M bbox points: the whole tree is cache resident
        z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
        v, f = z["elephant_v"], z["elephant_f"]
        g = B.AABB_tree(v, f)
        stream(g)
        g.build()
        bb = g.bbox()
        rays, pts = W.sphere_rays(bb, 1 << 20), W.bbox_points(bb, 1 << 20)
        rd, pd = torch.from_numpy(rays).cuda(), torch.from_numpy(pts).cuda()
        ms_r, _ = timed(lambda: g.first_intersection(rd))
        ms_p, _ = timed(lambda: g.closest_point_and_primitive(pd))
        r = {"triangles": len(f), "first_intersection_Mrays_s": len(rays) / ms_r / 1e3, "closest_point_Mqueries_s": len(pts) / ms_p / 1e3}
        if parity:
            o = orc.OracleTree(v, f)
            r["parity_rays_all"] = ray_parity(B, g, o, rays)
            r["parity_points_all"] = point_parity(B, g, o, pts)
        emit("config1_elephant", r)
        del g
    if "2" in which:      # icosphere level 8, 256^3 grid: unsigned distance field
        v, f = W.icosphere(8)
        g = B.AABB_tree(v, f)
        stream(g)
        g.build()
        pts_h = W.grid_points(256)
        pts = torch.from_numpy(pts_h).cuda()
        n = len(pts)
        ms, res = timed(lambda: g.closest_point_and_primitive(pts, want=("point", "sqdist")))
        st = g.stats()
        d = torch.sqrt(res["sqdist"])
        want = (torch.linalg.norm(pts, dim=1) - 1).abs()
        r = {"triangles": len(f), "points": n, "ms": ms, "closest_point_Mqueries_s": n / ms / 1e3, "build_ms": st["last_build_ms"],
             "nodes_per_query": st["nodes_visited"] / n, "leaves_per_query": st["leaves_tested"] / n, "max_abs_err_vs_unit_sphere": float((d - want).abs().max()),
             "ties": st["ties"]}
        assert r["max_abs_err_vs_unit_sphere"] < 2e-4     # unsigned distance field of the unit sphere, up to the facet sagitta
        if parity:
            o = orc.OracleTree(v, f)
            sub = pts_h[::16][:1 << 20]                                     # strided 1 M subset
            r["parity_strided_1M"] = point_parity(B, g, o, np.ascontiguousarray(sub))
            zs = np.concatenate([pts_h[k * 65536:(k + 1) * 65536] for k in (0, 128, 255)])   # three full z-slices (bottom, middle, top)
            r["parity_z_slices_0_128_255"] = point_parity(B, g, o, np.ascontiguousarray(zs))
        emit("config2_icosphere", r)
        del g, pts, res
    if "3" in which:      # terrain16M: primary rays + shadow segments
        v, f = W.terrain(4096, 2048)
        g = B.AABB_tree(v, f)
        stream(g)
        g.build()
        segs_h = W.shadow_segments(v, 1 << 24)
        segs = torch.from_numpy(segs_h).cuda()
        ms, hit = timed(lambda: g.do_intersect(segs, B.SEGMENT))
        st = g.stats()
        ms2, cnt = timed(lambda: g.number_of_intersected_primitives(segs, B.SEGMENT))
        st2 = g.stats()
        hit, cnt = hit.cpu().numpy(), cnt.cpu().numpy()
        assert np.array_equal(hit != 0, cnt > 0)
        r = {"triangles": len(f), "shadow_segments": len(segs), "do_intersect_ms": ms, "do_intersect_Mq_s": len(segs) / ms / 1e3, "in_shadow": float((hit != 0).mean()),
             "nodes_per_query": st["nodes_visited"] / len(segs), "leaves_per_query": st["leaves_tested"] / len(segs), "count_ms": ms2, "count_Mq_s": len(segs) / ms2 / 1e3,
             "count_nodes_per_query": st2["nodes_visited"] / len(segs), "count_leaves_per_query": st2["leaves_tested"] / len(segs), "exact_stage_queries": st2["exact_fallbacks"]}
        if parity:
            o = orc.OracleTree(v, f)
            rows = np.arange(0, 8192, 64)                                    # 128 full image rows = 1,048,576 primary rays
            rays = np.concatenate([W.pinhole_rays(8192, 8192, 4096, 2048, rows=(int(k), int(k) + 1)) for k in rows])
            r["parity_primary_rays_1M"] = ray_parity(B, g, o, rays)
            s1 = segs_h[::16]                                                # 1 M of the 16 M shadow segments
            want = o.count(s1, 0)
            r["parity_shadow_segments_1M"] = {"segments": len(s1), "count_mismatch": int((cnt[::16] != want).sum()), "do_intersect_mismatch": int(((hit[::16] != 0) != (want > 0)).sum())}
        emit("config3_terrain", r)
        del g, segs
    if "4" in which:      # random soup: build throughput + counts
        n = int(os.environ.get("SOUP_N", "100000000"))
        t0 = time.time()
        soup = W.triangle_soup(n)
        g = B.AABB_tree(triangles=soup)
        stream(g)
        g.build()
        bt = []
        for _ in range(3):
            g.build()
            bt.append(g.stats()["last_build_ms"])
        segs_h = W.short_segments(1 << 24)
        segs = torch.from_numpy(segs_h).cuda()
        ms, cnt = timed(lambda: g.number_of_intersected_primitives(segs, B.SEGMENT))
        ms2, hit = timed(lambda: g.do_intersect(segs, B.SEGMENT))
        cnt, hit = cnt.cpu().numpy(), hit.cpu().numpy()
        assert np.array_equal(hit != 0, cnt > 0)
        r = {"triangles": n, "host_setup_s": time.time() - t0, "build_ms": min(bt), "build_Mtri_s": n / min(bt) / 1e3, "build_frac_of_hbm_roofline(492B/tri)": 492.0 * n / min(bt) / 1e6 / 6551.4,
             "segments": len(segs), "count_ms": ms, "count_Mq_s": len(segs) / ms / 1e3, "do_intersect_ms": ms2, "mean_count": float(cnt.mean()), "max_count": int(cnt.max())}
        del g
        if parity:       # the 10 M-triangle prefix (the oracle at full size only if host RAM allows), 1 M segments
            m = min(n, 10_000_000)
            g2 = B.AABB_tree(triangles=soup[:m])
            stream(g2)
            o = orc.OracleTree(soup=soup[:m])
            s1 = np.ascontiguousarray(segs_h[:1 << 20])
            got = g2.number_of_intersected_primitives(s1, B.SEGMENT)
            want = o.count(s1, 0)
            r["parity_10M_prefix"] = {"triangles": m, "segments": len(s1), "count_mismatch": int((got != want).sum()), "total_intersections": int(want.sum()),
                                      "do_intersect_mismatch": int(((g2.do_intersect(s1, B.SEGMENT) != 0) != (want > 0)).sum())}
        emit("config4_soup", r)
    if out_path:
        json.dump(results, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
