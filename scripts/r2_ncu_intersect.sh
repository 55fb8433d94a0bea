#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:intersect_kernel --launch-skip 2 -c 4 -o gpurun_out/p_intersect -f python scripts/run_configs.py 3 --no-parity > gpurun_out/p_intersect.log 2>&1
tail -2 gpurun_out/p_intersect.log | cut -c1-300
ls -la gpurun_out/p_intersect*
