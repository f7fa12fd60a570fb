#!/bin/bash
# round 2, first GPU call: parity of the dense per-frame pass + the three lines VERDICT "Next 1" names
OUT=gpurun_out/r2a
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
for wl in c3-outlier-abs-extreme a4-gauss-noise a1-iid-uniform c4-outlier-rel-forward; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$wl.json"))
r=d["roofline"]
print("$wl", "ms/step %.3f"%d["ms_per_step"], "frac %.3f"%r["frac"], "launch_ms %.3f"%r["avg_launch_ms"], "main %.3f"%r.get("dominant_kernel",{}).get("avg_launch_ms",0), "slow", r["slow_path_pixels_per_launch"])
PY
done
timeout 300 python tools/quick_time.py 4000 6000 200 2,4 > $OUT/quick_c3.txt 2>&1; tail -14 $OUT/quick_c3.txt
