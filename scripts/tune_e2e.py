#!/usr/bin/env python
"""GPU-side sweep of the HOST-pointer pipeline (chunk count) on the benchmark batch + raw pinned-copy bandwidth.  Not a bench value."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    import bench
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.build()
    n = 1 << 23
    rays_h = bench.shard_rays(0)
    pts_h = W.bbox_points(g.bbox(), n)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    rays_p, pts_p = pinned((n, 6), torch.float64), pinned((n, 3), torch.float64)
    rays_p[:], pts_p[:] = rays_h, pts_h
    out_r = {"prim": pinned(n, torch.int32).view(np.uint32), "t": pinned(n, torch.float64), "point": pinned((n, 3), torch.float64), "flags": pinned(n, torch.uint8)}
    out_p = {"point": pinned((n, 3), torch.float64), "prim": pinned(n, torch.int32).view(np.uint32), "sqdist": pinned(n, torch.float64), "flags": pinned(n, torch.uint8)}
    # raw copy bandwidth
    d = torch.empty(n * 6, dtype=torch.float64, device="cuda")
    src = torch.from_numpy(rays_p.reshape(-1))
    for name, fn in (("H2D", lambda: d.copy_(src, non_blocking=True)), ("D2H", lambda: src.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        print("%s pinned: %.1f GB/s" % (name, 3 * n * 48 / (time.perf_counter() - t0) / 1e9), flush=True)
    rays_p[:] = rays_h
    for parts in (2, 3, 6, 12):
        g.set_option(B.OPT_HOST_PARTS, parts)
        for which in ("rays", "points"):
            def step():
                if which == "rays":
                    g.first_intersection(rays_p, want=("prim", "t", "point", "flags"), out=out_r)
                else:
                    g.closest_point_and_primitive(pts_p, want=("point", "prim", "sqdist", "flags"), out=out_p)
            step()
            t0 = time.perf_counter()
            for _ in range(3):
                step()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            print("parts=%2d %-6s %.2f ms  (%.0f Mq/s)" % (parts, which, ms, n / ms / 1e3), flush=True)


if __name__ == "__main__":
    main()
