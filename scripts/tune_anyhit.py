#!/usr/bin/env python
"""GPU-side timing of do_intersect on config 3 (terrain16M shadow segments) and a 32M soup (short segments).  Not a bench value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import torch
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    from run_configs import timed
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    segs = torch.from_numpy(W.shadow_segments(v, 1 << 24)).cuda()
    ms, _ = timed(lambda: g.do_intersect(segs, B.SEGMENT))
    ms2, _ = timed(lambda: g.number_of_intersected_primitives(segs, B.SEGMENT))
    print("terrain16M shadow do_intersect %.2f ms (%.0f Mq/s)  count %.2f ms" % (ms, len(segs) / ms / 1e3, ms2), flush=True)
    del g
    s = B.AABB_tree(triangles=W.triangle_soup(32_000_000))
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    q = torch.from_numpy(W.short_segments(1 << 24)).cuda()
    ms, _ = timed(lambda: s.do_intersect(q, B.SEGMENT))
    ms2, _ = timed(lambda: s.number_of_intersected_primitives(q, B.SEGMENT))
    print("soup32M short do_intersect %.2f ms (%.0f Mq/s)  count %.2f ms" % (ms, len(q) / ms / 1e3, ms2), flush=True)


if __name__ == "__main__":
    main()
