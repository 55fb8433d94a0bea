#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_comm.py -m gpu -x -q -k "build or tree or degenerate or chain or refit or state or kats or comm" > gpurun_out/r2_build_tests.log 2>&1
tail -5 gpurun_out/r2_build_tests.log
python scripts/time_build.py 32000000 2>&1 | grep -v "^\[b200"
