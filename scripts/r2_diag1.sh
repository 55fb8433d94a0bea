#!/bin/bash
# round-2 diagnostic 1: baseline closest_point on the bench batch, the perfect-initial-bound floor, and an ncu source capture
mkdir -p gpurun_out
python scripts/tune_closest.py --skip-old --sweep 160 > gpurun_out/d1_base.log 2>&1
python scripts/tune_closest.py --skip-old --sweep 160 --two-pass > gpurun_out/d1_twopass.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:closest_point_warp_kernel --launch-skip 2 -c 1 -o gpurun_out/d1_warp -f python scripts/tune_closest.py --skip-old --sweep 160 --reps 3 > gpurun_out/d1_ncu.log 2>&1
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/d1_bench.json 2> gpurun_out/d1_bench.err
tail -3 gpurun_out/d1_base.log gpurun_out/d1_twopass.log; tail -c 600 gpurun_out/d1_bench.json
