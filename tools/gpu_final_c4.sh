#!/bin/bash
# c4 evidence of the final build: launch list and one ncu capture of outlier_exact_kernel
OUT=gpurun_out/${1:-finc4}
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c4.csv python tools/launch_times.py 2160 3840 1000 1 > /dev/null 2>&1
bash tools/gpu_ncu_exact_c4.sh ${1:-finc4} > $OUT/ncu.log 2>&1; head -30 $OUT/ncu_exact_c4.summary.txt
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches_c4.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:60], r[-1])
PY
