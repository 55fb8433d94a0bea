#!/bin/bash
# pool kernel: one or two groups of node fetches in flight per step, at 20 / 18 / 16 warps per SM (variants built on the box)
mkdir -p gpurun_out
for v in "1 20" "2 20" "2 18" "2 16"; do
  set -- $v
  B200AABB_NVCC_EXTRA="-DB2_WP_GROUPS=$1 -DB2_WP_WARPS_PER_SM=$2" python -m cgal_b200._build --force > gpurun_out/dg_build_$1_$2.log 2>&1
  grep -i "closest_point_warp" gpurun_out/dg_build_$1_$2.log | head -2
  timeout 600 python scripts/tune_closest.py --cfgs 1:160,1:96 --label "groups=$1 warps/SM=$2" > gpurun_out/dg_$1_$2.log 2>&1
  cat gpurun_out/dg_$1_$2.log
done
