#!/bin/bash
# experiment r2_30: resident warps per SM of the pool kernel between the measured 16 / 20 / 24: 18, 19, 22 (register budget 112 / 104 / 88)
mkdir -p gpurun_out
cp cgal_b200/libb200aabb.so /tmp/keep.so
timeout 600 python scripts/tune_closest.py --cfgs 1:160 --reps 5 --label "20 warps/SM (default)" 2>&1 | tail -1
for w in 18 19 22; do
  cp cgal_b200/libb200aabb_w$w.so cgal_b200/libb200aabb.so
  timeout 600 python scripts/tune_closest.py --cfgs 1:160 --reps 5 --label "$w warps/SM" 2>&1 | tail -2
done
cp /tmp/keep.so cgal_b200/libb200aabb.so
