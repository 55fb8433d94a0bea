#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d11_pytest.log 2>&1
timeout 600 python scripts/tune_closest.py --cfgs 0:0,1:160,1:128,1:96 --label "quantised 32-byte nodes in the pool kernel" > gpurun_out/d11_tune.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:closest_point_warp_kernel --launch-skip 2 -c 1 -o gpurun_out/d11_warp -f python scripts/tune_closest.py --cfgs 1:160 --reps 3 > gpurun_out/d11_ncu.log 2>&1
tail -5 gpurun_out/d11_pytest.log; cat gpurun_out/d11_tune.log
