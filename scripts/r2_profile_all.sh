#!/bin/bash
# final-state ncu captures of every kernel of the bench step and of the build (one launch each, --set full), + the launch list of a bench run
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:first_intersection_kernel --launch-skip 3 -c 1 -o gpurun_out/p_first -f python bench.py --no-cpu-baseline --no-e2e --no-config5 --steps 1 --warmup 3 > gpurun_out/p_first.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:closest_point_warp_kernel --launch-skip 3 -c 1 -o gpurun_out/p_pool -f python bench.py --no-cpu-baseline --no-e2e --no-config5 --steps 1 --warmup 3 > gpurun_out/p_pool.log 2>&1
ncu --set full --clock-control none -k "regex:scene_bounds|morton|rs_onesweep|rs_scan|build_local|hierarchy|refit_global|gather_mesh" --launch-skip 10 -c 10 -o gpurun_out/p_build -f python scripts/time_build.py 1000000 > gpurun_out/p_build.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches.csv python bench.py --no-cpu-baseline --no-e2e --no-config5 --steps 2 --warmup 3 > gpurun_out/p_launches.log 2>&1
ls -la gpurun_out/p_*
