#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "segment_primitives or cpp" > gpurun_out/r2_25_gpu_tests.log 2>&1
tail -12 gpurun_out/r2_25_gpu_tests.log
