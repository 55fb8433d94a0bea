#!/bin/bash
# gpu tests + one bench line (no config 5) on the current tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_check_tests.log 2>&1
tail -5 gpurun_out/r2_check_tests.log
timeout 900 python bench.py --no-config5 > gpurun_out/r2_check_bench.json 2> gpurun_out/r2_check_bench.err
tail -c 3000 gpurun_out/r2_check_bench.json
