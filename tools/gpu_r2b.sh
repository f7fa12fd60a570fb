#!/bin/bash
OUT=gpurun_out/r2b
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:outlier_exact_kernel -s 1 -c 1 -f -o $OUT/ncu_exact_a4 \
  python tools/launch_times.py 4000 6000 200 0 4 > $OUT/ncu_exact_a4.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 40 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
cat $OUT/ncu_exact_a4.summary.txt
