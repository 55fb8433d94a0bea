#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d6_pytest.log 2>&1
B200AABB_BUILD_TIMING=1 timeout 600 python scripts/time_build.py 1000000 > gpurun_out/d6_build.log 2>&1
( time timeout 1200 python bench.py --steps 10 --warmup 3 ) > gpurun_out/d6_bench.json 2> gpurun_out/d6_bench.err
tail -4 gpurun_out/d6_pytest.log; grep -v "^\[b200aabb build\] \(root\|scene\)" gpurun_out/d6_build.log | head -12; tail -5 gpurun_out/d6_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/d6_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); det=d['detail']
    for k in ('first_intersection_Mrays_s','first_intersection_hits_Mrays_s','first_intersection_all_hit_camera','closest_point_Mqueries_s','build_ms','fp64_dfma_peak_TFLOP_s'): print(k, det[k])
    c5=det['config5']; print({k:c5[k] for k in c5 if k!='checksum128_per_array'})
    print(d['cpu_baseline'])
except Exception as e: print('ERR',e); print(open('gpurun_out/d6_bench.json').read()[-2000:])
PY
