#!/bin/bash
TAG=${1:-ncu_exact_c4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^outlier_exact_kernel -s 1 -c 1 -f -o $OUT/ncu_exact_c4 \
  python tools/launch_times.py 2160 3840 1000 1 > $OUT/ncu_exact_c4.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 60 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
head -95 $OUT/ncu_exact_c4.summary.txt
