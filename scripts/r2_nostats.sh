#!/bin/bash
# experiment r2_29: pool kernel with the per-round traversal counters (w_nodes / w_leaves / w_evals) compiled out (-DB2_WP_NO_STATS)
mkdir -p gpurun_out
timeout 600 python scripts/tune_closest.py --cfgs 1:160 --reps 5 --label "with counters" 2>&1 | tail -2
cp cgal_b200/libb200aabb.so /tmp/keep.so; cp cgal_b200/libb200aabb_nostats.so cgal_b200/libb200aabb.so
timeout 600 python scripts/tune_closest.py --cfgs 1:160 --reps 5 --label "counters compiled out" 2>&1 | tail -2
cp /tmp/keep.so cgal_b200/libb200aabb.so
