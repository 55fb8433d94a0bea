#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k "regex:build_local|hierarchy|refit_global" --launch-skip 6 -c 3 -o gpurun_out/p_build2 -f python scripts/time_build.py 1000000 > gpurun_out/p_build2.log 2>&1
tail -3 gpurun_out/p_build2.log
ls -la gpurun_out/p_build2*
