#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_22_gpu_tests.log 2>&1
tail -15 gpurun_out/r2_22_gpu_tests.log
