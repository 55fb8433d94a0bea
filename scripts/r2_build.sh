#!/bin/bash
# fused bottom-up build: parity tests of the build, stage timing fused vs classic
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "build or tree or degenerate or chain or refit" > gpurun_out/r2_build_tests.log 2>&1
tail -15 gpurun_out/r2_build_tests.log
B200AABB_BUILD_TIMING=1 timeout 600 python scripts/time_build.py 32000000 > gpurun_out/r2_build_time.log 2>&1
grep -v "^\[b200aabb build\]" gpurun_out/r2_build_time.log | tail -4
grep "^\[b200aabb build\]" gpurun_out/r2_build_time.log | sed -n 40,52p
python - <<'PY' >> gpurun_out/r2_build_time.log 2>&1
import sys
sys.path.insert(0, '.')
import cgal_b200 as B
from cgal_b200 import workloads as W
v, f = W.terrain(4096, 2048)
g = B.AABB_tree(v, f)
g.set_option(103, 1)
g.build()
ts = []
for _ in range(5):
    g.build(); ts.append(g.stats()["last_build_ms"])
print("classic path terrain16M build ms:", ["%.3f" % t for t in ts])
PY
tail -1 gpurun_out/r2_build_time.log
