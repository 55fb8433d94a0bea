#!/usr/bin/env python
"""first_intersection on the benchmark batch (terrain16M, 8,388,608 pinhole rays, device resident): thread-per-ray kernel against the persistent-thread kernel with
ray replacement (B200AABB_OPT_RAY_KERNEL); checks that both return identical records.  Not a bench value."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import cgal_b200 as B
    from cgal_b200 import device_workloads as D
    from cgal_b200 import workloads as W
    from cgal_b200.aabb_tree import OPT_RAY_KERNEL
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.build()
    import bench
    rays = bench.device_shard_rays(0, 0, torch) if hasattr(bench, "device_shard_rays") else torch.from_numpy(bench.shard_rays(0)).cuda()
    n = len(rays)
    first = None
    for opt in (0, 1, 0, 1):
        g.set_option(OPT_RAY_KERNEL, opt)
        best = None
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); out = g.first_intersection(rays); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        st = g.stats()
        same = True if first is None else all(torch.equal(out[k], first[k]) for k in out)
        if first is None:
            first = {k: x.clone() for k, x in out.items()}
        print(json.dumps({"ray_kernel": opt, "ms_incl_sort": round(best, 3), "Mrays_s": round(n / best / 1e3, 1), "nodes_per_ray": st["nodes_visited"] / n,
                          "identical_to_first": same}), flush=True)


if __name__ == "__main__":
    main()
