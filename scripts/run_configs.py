#!/usr/bin/env python
"""BASELINE.json configs[1] and configs[3] at full size on one GPU (device-resident queries): throughput for the record
(profiles/README.md) + size-independent checks.  Not the bench line (bench.py measures configs[2]/[4])."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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


def main():
    import torch
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    which = sys.argv[1:] or ["2", "4"]
    if "2" in which:
        v, f = W.icosphere(8)
        g = B.AABB_tree(v, f)
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        g.build()
        pts = torch.from_numpy(W.grid_points(256)).cuda()
        n = len(pts)
        ms, r = timed(lambda: g.closest_point_and_primitive(pts, want=("point", "sqdist")))
        st = g.stats()
        d = torch.sqrt(r["sqdist"])
        want = (torch.linalg.norm(pts, dim=1) - 1).abs()
        print(json.dumps({"config": 2, "triangles": len(f), "points": n, "ms": round(ms, 2), "Mq_s": round(n / ms / 1e3, 1), "build_ms": round(st["last_build_ms"], 3),
                          "nodes_q": round(st["nodes_visited"] / n, 1), "leaves_q": round(st["leaves_tested"] / n, 1),
                          "max_abs_err_vs_sphere": float((d - want).abs().max()), "ties": st["ties"]}), flush=True)
        assert float((d - want).abs().max()) < 2e-4          # unsigned distance field of the unit sphere, up to the facet sagitta
        del g, pts, r
    if "3" in which:
        v, f = W.terrain(4096, 2048)
        g = B.AABB_tree(v, f)
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        g.build()
        segs = torch.from_numpy(W.shadow_segments(v, 1 << 24)).cuda()
        ms, hit = timed(lambda: g.do_intersect(segs, B.SEGMENT))
        st = g.stats()
        ms2, cnt = timed(lambda: g.number_of_intersected_primitives(segs, B.SEGMENT))
        st2 = g.stats()
        hit, cnt = hit.cpu().numpy(), cnt.cpu().numpy()
        assert np.array_equal(hit != 0, cnt > 0)
        print(json.dumps({"config": 3, "triangles": len(f), "shadow_segments": len(segs), "do_intersect_ms": round(ms, 2), "do_intersect_Mq_s": round(len(segs) / ms / 1e3, 1),
                          "in_shadow": float((hit != 0).mean()), "nodes_q": round(st["nodes_visited"] / len(segs), 1), "leaves_q": round(st["leaves_tested"] / len(segs), 1),
                          "count_ms": round(ms2, 2), "count_nodes_q": round(st2["nodes_visited"] / len(segs), 1), "count_leaves_q": round(st2["leaves_tested"] / len(segs), 1),
                          "exact_stage_queries": st2["exact_fallbacks"]}), flush=True)
        del g, segs
    if "4" in which:
        n = int(os.environ.get("SOUP_N", "100000000"))
        t0 = time.time()
        soup = W.triangle_soup(n)
        g = B.AABB_tree(triangles=soup)
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        g.build()
        bt = []
        for _ in range(3):
            g.build()
            bt.append(g.stats()["last_build_ms"])
        segs = torch.from_numpy(W.short_segments(1 << 24)).cuda()
        ms, cnt = timed(lambda: g.number_of_intersected_primitives(segs, B.SEGMENT))
        ms2, hit = timed(lambda: g.do_intersect(segs, B.SEGMENT))
        cnt, hit = cnt.cpu().numpy(), hit.cpu().numpy()
        assert np.array_equal(hit != 0, cnt > 0)
        print(json.dumps({"config": 4, "triangles": n, "host_setup_s": round(time.time() - t0, 1), "build_ms": round(min(bt), 3), "build_Mtri_s": round(n / min(bt) / 1e3, 1),
                          "segments": len(segs), "count_ms": round(ms, 2), "count_Mq_s": round(len(segs) / ms / 1e3, 1), "do_intersect_ms": round(ms2, 2),
                          "mean_count": float(cnt.mean()), "max_count": int(cnt.max())}), flush=True)


if __name__ == "__main__":
    main()
