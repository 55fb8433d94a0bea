#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "side_of_mesh or line_and_plane or cpp or fused" > gpurun_out/r2_new_tests.log 2>&1
tail -25 gpurun_out/r2_new_tests.log
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import cgal_b200 as B
from cgal_b200 import workloads as W
v, f = W.icosphere(8)
g = B.AABB_tree(v, f); g.build()
pts = torch.from_numpy(np.random.default_rng(0).uniform(-1.05, 1.05, (1 << 23, 3))).cuda()
g.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); s = g.side_of_mesh(pts); e1.record(); torch.cuda.synchronize()
    print("side_of_mesh icosphere(8) 1.3M triangles, 8.4M points: %.2f ms = %.0f Mq/s; inside %.3f" % (e0.elapsed_time(e1), (1 << 23) / e0.elapsed_time(e1) / 1e3, float((s == 1).float().mean())))
PY
