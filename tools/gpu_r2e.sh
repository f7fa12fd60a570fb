#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
for wl in c3-outlier-abs-extreme a4-gauss-noise a1-iid-uniform; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e --no-verify > $OUT/bench_${wl}.json 2> $OUT/bench_${wl}.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_${wl}.json"))
r=d["roofline"]
print("$wl", "ms/step %.3f"%d["ms_per_step"], "frac %.3f"%r["frac"], "launch_ms %.3f"%r["avg_launch_ms"], "main %.3f"%r.get("dominant_kernel",{}).get("avg_launch_ms",0), "slow", r["slow_path_pixels_per_launch"])
PY
  tail -3 $OUT/bench_${wl}.err
done
echo "=== default bench (all workloads)"
( time timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | grep real
tail -30 $OUT/bench_default.err
python - <<PY
import json
d=json.load(open("$OUT/bench_default.json"))
def show(o):
    r=o.get("roofline") or {}
    print(o["config"]["workload"], "value %.3g"%o.get("value",0), "ms %.3f"%o.get("ms_per_step",0), "frac %.3f"%r.get("frac",0), "e2e", (o.get("e2e") or {}).get("value"), "cpu", (o.get("cpu_baseline") or {}).get("value"), "verified", o.get("verified"), o.get("error"))
show(d)
for o in d.get("other_workloads",[]): show(o)
print(d.get("other_workloads_skipped"))
PY
echo "=== reference arm"
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err ) 2>&1 | grep real
cut -c1-400 $OUT/bench_reference.json
