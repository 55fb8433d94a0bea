#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d5_pytest.log 2>&1
B200AABB_BUILD_TIMING=1 timeout 600 python scripts/time_build.py > gpurun_out/d5_build.log 2>&1
tail -15 gpurun_out/d5_pytest.log; tail -40 gpurun_out/d5_build.log
