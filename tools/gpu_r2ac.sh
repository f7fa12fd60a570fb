#!/bin/bash
OUT=gpurun_out/r2ac
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $OUT/racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|sanitize run done" $OUT/racecheck.txt; grep -c "Race reported" $OUT/racecheck.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $OUT/memcheck.txt 2>&1
grep -E "ERROR SUMMARY|sanitize run done" $OUT/memcheck.txt
timeout 1200 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; tail -2 $OUT/pytest_gpu.txt
for wl in c3-outlier-abs-extreme a4-gauss-noise a1-iid-uniform; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-others 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=o.get('roofline') or {}
print('$wl ms %.4f call %.4f frac %.3f tiers %.3f verified %s' % (o['ms_per_step'], r.get('avg_launch_ms') or 0, r.get('frac',0), r.get('tier_kernels_ms') or 0, (o.get('verified') or {}).get('ok')))"
done
