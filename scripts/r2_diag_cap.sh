#!/bin/bash
# pool kernel: size of the shared-memory node pool (what is left of the 256 KB is L1) and resident warps per SM
mkdir -p gpurun_out
for v in "768 20" "512 20" "384 20" "512 24" "384 24" "384 28"; do
  set -- $v
  B200AABB_NVCC_EXTRA="-DB2_WP_NODE_CAP=$1 -DB2_WP_WARPS_PER_SM=$2" python -m cgal_b200._build --force > gpurun_out/dc_build_$1_$2.log 2>&1
  grep -i "closest_point_warp" gpurun_out/dc_build_$1_$2.log | head -1
  timeout 600 python scripts/tune_closest.py --cfgs 1:160 --label "node pool $1 entries, warps/SM=$2" > gpurun_out/dc_$1_$2.log 2>&1
  cat gpurun_out/dc_$1_$2.log
done
