#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python scripts/run_configs.py 1 2 3 4 --out gpurun_out/r2_parity_configs.json > gpurun_out/r2_parity.log 2>&1
tail -c 1500 gpurun_out/r2_parity.log
