#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d8_pytest.log 2>&1
STEPS=6 timeout 600 python scripts/tune_e2e.py > gpurun_out/d8_e2e.log 2>&1
tail -6 gpurun_out/d8_pytest.log; cat gpurun_out/d8_e2e.log
