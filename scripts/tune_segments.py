#!/usr/bin/env python
"""do_intersect / number_of_intersected_primitives of the config 3 shadow segments (terrain16M, 16,777,216 segments, device resident): thread-per-query kernel
against the persistent-thread kernel with query replacement (B200AABB_OPT_RAY_KERNEL); identical results required.  Not a bench value."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    from cgal_b200.aabb_tree import OPT_RAY_KERNEL
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.build()
    segs = torch.from_numpy(W.shadow_segments(v, 1 << 24)).cuda()
    n = len(segs)
    first = {}
    for opt in (0, 1, 0, 1):
        g.set_option(OPT_RAY_KERNEL, opt)
        for name, fn in (("do_intersect", lambda: g.do_intersect(segs, B.SEGMENT)), ("count", lambda: g.number_of_intersected_primitives(segs, B.SEGMENT))):
            best = None
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            st = g.stats()
            same = torch.equal(out, first[name]) if name in first else True
            first.setdefault(name, out.clone())
            print(json.dumps({"kernel": opt, "query": name, "ms_incl_sort": round(best, 3), "Mseg_s": round(n / best / 1e3, 1), "nodes_per_query": round(st["nodes_visited"] / n, 2),
                              "identical_to_first": same}), flush=True)


if __name__ == "__main__":
    main()
