#!/usr/bin/env python
"""GPU-side tuning sweep for closest_point on the benchmark workload (terrain16M + bbox points): compares the closest_point kernels
(0 thread per query, 1 warp-cooperative pool, 2 lane pairs with private stacks) over their scheduling parameter and checks that they
return identical records.  Not a bench value."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=2048)
    ap.add_argument("--points", type=int, default=1 << 23)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cfgs", default="1:160,2:64", help="comma separated kernel:exact_go")
    ap.add_argument("--label", default="")
    args = ap.parse_args()
    import torch
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    t0 = time.time()
    v, f = W.terrain(args.nx, args.ny)
    g = B.AABB_tree(v, f)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.build()
    pts = torch.from_numpy(W.bbox_points(g.bbox(), args.points)).cuda()
    n = args.points
    out = {"point": torch.empty((n, 3), dtype=torch.float64, device="cuda"), "prim": torch.empty(n, dtype=torch.int32, device="cuda"),
           "sqdist": torch.empty(n, dtype=torch.float64, device="cuda"), "flags": torch.empty(n, dtype=torch.uint8, device="cuda")}
    want = ("point", "prim", "sqdist", "flags")
    print("setup %.1fs, %d triangles, %d points %s" % (time.time() - t0, len(f), n, args.label), flush=True)

    def run(label):
        for k in out:
            out[k].zero_()
        best = None
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            g.closest_point_and_primitive(pts, want=want, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        st = g.stats()
        print(json.dumps({"cfg": label, "ms": round(best, 3), "Mq_s": round(n / best / 1e3, 1), "nodes_q": round(st["nodes_visited"] / n, 1),
                          "leaves_q": round(st["leaves_tested"] / n, 1), "culled_q": round(st["leaves_culled"] / n, 1), "ties": st["ties"], "exact": st["exact_fallbacks"]}), flush=True)
        return {k: out[k].clone() for k in ("prim", "sqdist", "point")}, out["flags"].clone()

    ref = None
    for item in args.cfgs.split(","):
        kern, go = [int(x) for x in item.split(":")]
        g.set_option(B.OPT_CLOSEST_KERNEL, kern)
        g.set_option(B.OPT_WARP_EXACT_GO, go)
        got, fl = run("kernel=%d exact_go=%d" % (kern, go))
        if ref is None:
            ref, ref_fl = got, fl
        else:
            same = all(bool(torch.equal(got[k], ref[k])) for k in got) and bool(torch.equal(fl & B.FLAG_TIE, ref_fl & B.FLAG_TIE))
            print("   identical to first result:", same, flush=True)


if __name__ == "__main__":
    main()
