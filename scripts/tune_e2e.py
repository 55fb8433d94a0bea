#!/usr/bin/env python
"""GPU-side look at the HOST-pointer pipeline on the benchmark batch: asynchronous steps (B200AABB_OPT_ASYNC_HOST) over chunk counts, and —
with B200AABB_PIPE_TRACE=1 — the per-chunk timeline (upload / kernels / download) the library prints at b200aabb_synchronize.
Not a bench value."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import cgal_b200 as B
    from cgal_b200 import device_workloads as D
    from cgal_b200 import workloads as W
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    g.build()
    n = 1 << 23
    rays_d = torch.empty((n, 6), dtype=torch.float64, device="cuda")
    cam = D.pinhole_camera(4096, 2048)
    for j in range(1024):
        D.pinhole_rays(cam, 8192, 8192, j * 8, 1, rays_d[j * 8192:(j + 1) * 8192])
    pts_d = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    D.bbox_points(g.bbox(), n, pts_d)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    rays_p, pts_p = pinned((n, 6), torch.float64), pinned((n, 3), torch.float64)
    rays_p[:], pts_p[:] = rays_d.cpu().numpy(), pts_d.cpu().numpy()
    out_r = {"prim": pinned(n, torch.int32).view(np.uint32), "t": pinned(n, torch.float64), "point": pinned((n, 3), torch.float64), "flags": pinned(n, torch.uint8)}
    out_p = {"point": pinned((n, 3), torch.float64), "prim": pinned(n, torch.int32).view(np.uint32), "sqdist": pinned(n, torch.float64), "flags": pinned(n, torch.uint8)}
    want_r, want_p = ("prim", "t", "point", "flags"), ("point", "prim", "sqdist", "flags")
    steps = int(os.environ.get("STEPS", "4"))
    g.set_option(B.OPT_ASYNC_HOST, 1)
    for rparts, pparts in [(8, 1)] if os.environ.get("B200AABB_PIPE_TRACE") else [(8, 1), (4, 1), (16, 1), (8, 2)]:
        def step():
            g.set_option(B.OPT_HOST_PARTS, rparts)
            g.first_intersection(rays_p, want=want_r, out=out_r)
            g.set_option(B.OPT_HOST_PARTS, pparts)
            g.closest_point_and_primitive(pts_p, want=want_p, out=out_p)
        step()
        g.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        g.synchronize()
        ms = (time.perf_counter() - t0) / steps * 1e3
        print("ray chunks %2d point chunks %2d: %.2f ms per step (%.0f Mq/s)" % (rparts, pparts, ms, 2 * n / ms / 1e3), flush=True)


if __name__ == "__main__":
    main()
