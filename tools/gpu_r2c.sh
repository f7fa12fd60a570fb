#!/bin/bash
OUT=gpurun_out/r2c
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
for im in 0 8 12 16 24; do
for wl in c3-outlier-abs-extreme a4-gauss-noise a1-iid-uniform; do
  CHB_INLINE_MIN=$im timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e > $OUT/bench_${wl}_$im.json 2> $OUT/bench_${wl}_$im.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_${wl}_$im.json"))
r=d["roofline"]
print("inline_min $im", "$wl", "ms/step %.3f"%d["ms_per_step"], "frac %.3f"%r["frac"], "launch_ms %.3f"%r["avg_launch_ms"], "main %.3f"%r.get("dominant_kernel",{}).get("avg_launch_ms",0), "slow", r["slow_path_pixels_per_launch"])
PY
done
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_main_a4 \
  python tools/launch_times.py 4000 6000 200 0 4 > $OUT/ncu_main_a4.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 40 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
cat $OUT/ncu_main_a4.summary.txt
