#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/dbg_side.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import cgal_b200 as B
z = np.load('tests/golden/meshes.npz')
v, f = z['elephant_v'], z['elephant_f']
g = B.AABB_tree(v, f); g.build()
pts = np.random.default_rng(0).uniform(v.min(0), v.max(0), (5000, 3))
a = g.side_of_mesh(pts)
print("host ok", (a == 1).mean())
b, att = g.side_of_mesh(torch.from_numpy(pts).cuda(), want_attempts=True)
torch.cuda.synchronize()
print("device+attempts ok", np.array_equal(b.cpu().numpy(), a))
c = g.side_of_mesh(torch.from_numpy(pts).cuda())
torch.cuda.synchronize()
print("device ok", np.array_equal(c.cpu().numpy(), a))
PY
compute-sanitizer --tool memcheck python /tmp/dbg_side.py > gpurun_out/dbg_side.log 2>&1
grep -v "^=========     at\|^=========     by\|Host Frame\|^=========$" gpurun_out/dbg_side.log | head -40
