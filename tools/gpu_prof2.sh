#!/bin/bash
# Second profiling pass: the tier kernels (launch list at c4, ncu --set full of outlier_exact_kernel at c4 and c3, outlier_hard_kernel at c3).
TAG=${1:-r1c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c4.csv python tools/launch_times.py 2160 3840 1000 1 > $OUT/launches_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:outlier_exact_kernel -s 1 -c 1 -f -o $OUT/ncu_exact_c4 python tools/launch_times.py 2160 3840 1000 1 > $OUT/ncu_exact_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:outlier_exact_kernel -s 1 -c 1 -f -o $OUT/ncu_exact_c3 python tools/launch_times.py 4000 6000 200 0 > $OUT/ncu_exact_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:outlier_hard_kernel -s 1 -c 1 -f -o $OUT/ncu_hard_c3 python tools/launch_times.py 4000 6000 200 0 > $OUT/ncu_hard_c3.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 45 > $b.summary.txt 2>&1
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  rm -f $r
done
timeout 200 python tools/quick_time.py 2160 3840 1000 2 > $OUT/quick_c4.txt 2>&1
cat $OUT/quick_c4.txt
ls -la $OUT
