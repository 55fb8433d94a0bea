#!/bin/bash
# one-GPU records on the state after the Triangle_3 / tree-vs-tree work: bench line, reference arm
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_24_bench_1gpu.json 2> gpurun_out/r2_24_bench_1gpu.err
tail -c 1800 gpurun_out/r2_24_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_24_reference_arm.json 2> gpurun_out/r2_24_reference_arm.err
tail -c 500 gpurun_out/r2_24_reference_arm.json
