#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python scripts/run_configs.py 3 4 --out gpurun_out/r2_parity_configs_3_4_fused.json > gpurun_out/r2_cfg34.log 2>&1
tail -c 3000 gpurun_out/r2_cfg34.log
