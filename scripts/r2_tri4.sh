#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "transformations or tree_against_tree or cpp" 2>&1 | tail -5
timeout 900 python scripts/time_triangle_queries.py --out gpurun_out/r2_21_triangle_queries.json 2>&1 | tail -12
