#!/bin/bash
# quick iteration: tests, a1 / c3 / c4 / c5 kernel-only bench lines, two shard timings.   bash tools/gpu_iter_quick.sh [tag]
TAG=${1:-r2w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
for wl in c3-outlier-abs-extreme a1-iid-uniform a4-gauss-noise c4-outlier-rel-forward c5-video; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-others > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  python - <<PY
import json
o=json.load(open("$OUT/bench_$wl.json"))
r=o.get("roofline") or {}; v=o.get("verified") or {}; dk=(r.get("dominant_kernel") or {})
print(o["config"]["workload"], "ms %.3f"%o.get("ms_per_step",0), "call %.3f"%(r.get("avg_launch_ms") or 0), "frac %.3f"%r.get("frac",0), "main %.3f"%(dk.get("avg_launch_ms") or 0), "tiers %.3f"%(r.get("tier_kernels_ms") or 0), "verified", v.get("ok"), v.get("pixels_differing_from_oracle"))
PY
done
timeout 300 python tools/small_band.py 2>&1 | grep -v "^rows 2000\|^rows 1000\|G=2\|G=4" | tee $OUT/small_band.txt
