#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d10_pytest.log 2>&1
STEPS=6 timeout 600 python scripts/tune_e2e.py > gpurun_out/d10_e2e.log 2>&1
STEPS=3 B200AABB_PIPE_TRACE=1 timeout 600 python scripts/tune_e2e.py > gpurun_out/d10_trace.log 2>&1
tail -6 gpurun_out/d10_pytest.log; cat gpurun_out/d10_e2e.log; tail -30 gpurun_out/d10_trace.log
