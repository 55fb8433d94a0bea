#!/bin/bash
mkdir -p gpurun_out
STEPS=3 B200AABB_PIPE_TRACE=1 timeout 600 python scripts/tune_e2e.py > gpurun_out/d9_trace.log 2>&1
grep -c pipe gpurun_out/d9_trace.log; tail -62 gpurun_out/d9_trace.log
