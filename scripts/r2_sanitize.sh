#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1
tail -4 gpurun_out/r2_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py > gpurun_out/r2_racecheck.log 2>&1
grep -c "Race reported" gpurun_out/r2_racecheck.log
grep "Race reported\|RACECHECK SUMMARY\|^ok" gpurun_out/r2_racecheck.log | sed 's/=========//' | sort | uniq -c | sort -rn | head -20
