#!/bin/bash
OUT=gpurun_out/r2d
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
for k in 3 4; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_kind$k.csv python tools/launch_times.py $([ $k = 3 ] && echo "2048 2048" || echo "4000 6000") 200 0 $k > $OUT/launches_kind$k.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches_kind$k.csv")) if len(r)>5 and r[0].isdigit()]
from collections import defaultdict
t=defaultdict(list)
for r in rows: t[r[4].split('(')[0][:60]].append(float(r[-1].replace(',','')))
for k_,v in t.items(): print("kind $k", k_, ["%.1f"%(x/1e3) for x in v[-4:]])
PY
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_main_a4 \
  python tools/launch_times.py 4000 6000 200 0 4 > $OUT/ncu_main_a4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:outlier_hard_kernel -s 1 -c 1 -f -o $OUT/ncu_hard_a1 \
  python tools/launch_times.py 2048 2048 200 0 3 > $OUT/ncu_hard_a1.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 30 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
cat $OUT/ncu_main_a4.summary.txt $OUT/ncu_hard_a1.summary.txt
