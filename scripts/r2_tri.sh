#!/bin/bash
# Triangle_3 queries + tree-vs-tree do_intersect: the new GPU tests, the C++ host test, the timing script
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "triangle_queries or tree_against_tree or cpp" > gpurun_out/r2_tri_tests.log 2>&1
tail -15 gpurun_out/r2_tri_tests.log
timeout 900 python scripts/time_triangle_queries.py --out gpurun_out/r2_21_triangle_queries.json 2>&1 | tail -20
