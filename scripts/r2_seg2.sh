#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "segment or closest or config1 or cpp" > gpurun_out/r2_28_gpu_tests.log 2>&1
tail -4 gpurun_out/r2_28_gpu_tests.log
timeout 900 python scripts/time_triangle_queries.py --out gpurun_out/r2_21_triangle_queries.json 2>&1 | grep "segments"
timeout 1500 python bench.py > gpurun_out/r2_28_bench_1gpu.json 2> gpurun_out/r2_28_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_28_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['detail']['closest_point_Mqueries_s'], d['detail']['first_intersection_Mrays_s'], d['detail']['build_ms'])
PY
