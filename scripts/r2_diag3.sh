#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/tune_closest.py --cfgs 1:160,2:32,2:64,2:128 --label "20 warps/SM" > gpurun_out/d3_w20.log 2>&1
for wps in 16 24; do
  B200AABB_NVCC_EXTRA="-DB2_PP_WARPS_PER_SM=$wps" python -m cgal_b200._build --force > gpurun_out/d3_build_$wps.log 2>&1
  timeout 600 python scripts/tune_closest.py --cfgs 2:64 --label "$wps warps/SM" > gpurun_out/d3_w$wps.log 2>&1
done
python -m cgal_b200._build --force > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:closest_point_pair_kernel --launch-skip 2 -c 1 -o gpurun_out/d3_pair -f python scripts/tune_closest.py --cfgs 2:64 --reps 3 > gpurun_out/d3_ncu.log 2>&1
cat gpurun_out/d3_w20.log gpurun_out/d3_w16.log gpurun_out/d3_w24.log
