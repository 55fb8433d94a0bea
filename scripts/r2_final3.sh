#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_32_bench_1gpu.json 2> gpurun_out/r2_32_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_32_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['detail']['closest_point_Mqueries_s'], d['detail']['first_intersection_Mrays_s'], d['detail']['build_ms'], d['gpu_launches'])
print(json.dumps(d['detail']['other_queries']))
print(d['cpu_baseline']['parity_on_sample'] if 'parity_on_sample' in d['cpu_baseline'] else '')
PY
tail -3 gpurun_out/r2_32_bench_1gpu.err
