#!/bin/bash
# C++ host test (collision-detection consumer, other primitive types), memcheck over the small run incl. the triangle kernels, ncu of the triangle kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "cpp or transformations" 2>&1 | tail -4
timeout 1200 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r2_23_memcheck.log 2>&1
tail -5 gpurun_out/r2_23_memcheck.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"triangle_kernel|tree_tree_kernel" -c 6 -o gpurun_out/p_triangle -f python scripts/time_triangle_queries.py > gpurun_out/p_triangle.log 2>&1
tail -3 gpurun_out/p_triangle.log | cut -c1-200
ls -la gpurun_out/p_triangle*
