#!/bin/bash
mkdir -p gpurun_out
for b in 8 6 5; do
  B200AABB_NVCC_EXTRA="-DB2_IS_BLOCKS=$b" python -m cgal_b200._build --force > gpurun_out/ds_build_$b.log 2>&1
  echo "blocks per SM $b"
  timeout 600 python scripts/tune_segments.py 2>&1 | tail -8
done
