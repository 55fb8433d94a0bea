#!/bin/bash
# final single-GPU records of the round: bench line (with config 5), reference arm, gpu tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1
tail -3 gpurun_out/r2_final_tests.log
timeout 1500 python bench.py > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
tail -c 1500 gpurun_out/r2_final_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err
tail -c 600 gpurun_out/r2_final_reference_arm.json
