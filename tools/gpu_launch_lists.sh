#!/bin/bash
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^outlier_hard_kernel -s 1 -c 1 -f -o $OUT/ncu_hard_a1 \
  python tools/launch_times.py 2048 2048 200 0 3 > $OUT/ncu_hard_a1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_a1.csv python tools/launch_times.py 2048 2048 200 0 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c4.csv python tools/launch_times.py 2160 3840 1000 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches_c5.csv python tools/prof_c5_ncu.py > /dev/null 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 40 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
head -70 $OUT/ncu_hard_a1.summary.txt
for f in a1 c4 c5; do echo == $f; python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches_$f.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-12:]: print(r[4][:60], r[-1])
PY
done
