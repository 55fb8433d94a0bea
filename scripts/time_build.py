#!/usr/bin/env python
"""GPU-side build timing (terrain16M and a 32M soup): device time of b200aabb_build, best of 5.  Not a bench value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cgal_b200 as B
    from cgal_b200 import workloads as W
    v, f = W.terrain(4096, 2048)
    g = B.AABB_tree(v, f)
    g.build()
    ts = []
    for _ in range(5):
        g.build()
        ts.append(g.stats()["last_build_ms"])
    print("terrain16M build ms:", ["%.3f" % t for t in ts], "-> %.0f Mtri/s" % (len(f) / min(ts) / 1e3), flush=True)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
    s = B.AABB_tree(triangles=W.triangle_soup(n))
    s.build()
    ts = []
    for _ in range(3):
        s.build()
        ts.append(s.stats()["last_build_ms"])
    print("soup %d build ms:" % n, ["%.3f" % t for t in ts], "-> %.0f Mtri/s" % (n / min(ts) / 1e3), flush=True)


if __name__ == "__main__":
    main()
