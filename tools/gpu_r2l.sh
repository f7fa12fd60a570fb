#!/bin/bash
# A/B: cp.async ring in the dense pass, bracket between the solver's two jumps, direct-evaluation video kernel.
OUT=gpurun_out/r2n
mkdir -p $OUT
export PYTHONUNBUFFERED=1
show() { python -c "
import sys,json
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; v=d.get('verified') or {}
print('$2', d['config']['workload'], 'ms %.3f'%d['ms_per_step'], 'call %.3f'%r.get('avg_launch_ms',0), 'frac %.3f'%r['frac'], 'main %.3f'%(r.get('dominant_kernel') or {}).get('avg_launch_ms',0), 'verified', v.get('ok'), v.get('pixels_differing_from_oracle'))
"; }
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
for wl in a4-gauss-noise a1-iid-uniform c3-outlier-abs-extreme c1-minimal; do
  CHB_LIB=$PWD/chrono_photo_b200/_variants/base.so timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --no-verify --steps 10 --warmup 3 > $OUT/base_$wl.json 2> $OUT/base_$wl.err; show $OUT/base_$wl.json base
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 10 --warmup 3 > $OUT/new_$wl.json 2> $OUT/new_$wl.err; show $OUT/new_$wl.json new
done
CHB_VIDEO_DIRECT=0 timeout 400 python bench.py --workload c5-video --no-cpu --no-e2e --no-verify --steps 10 --warmup 3 > $OUT/slide_c5.json 2> $OUT/slide_c5.err; show $OUT/slide_c5.json sliding
timeout 400 python bench.py --workload c5-video --no-cpu --no-e2e --steps 10 --warmup 3 > $OUT/direct_c5.json 2> $OUT/direct_c5.err; show $OUT/direct_c5.json direct
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^video_direct_kernel -s 1 -c 1 -f -o $OUT/ncu_video_c5 python tools/prof_c5_ncu.py > $OUT/ncu_video_c5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^outlier_hard_kernel -s 1 -c 1 -f -o $OUT/ncu_hard_a1 python tools/launch_times.py 2048 2048 200 0 3 > $OUT/ncu_hard_a1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_outlier_a4 python tools/launch_times.py 4000 6000 200 0 4 > $OUT/ncu_outlier_a4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_a1.csv python tools/launch_times.py 2048 2048 200 0 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c3.csv python tools/launch_times.py 4000 6000 200 0 2 > /dev/null 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 30 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_band500.csv python tools/small_band_once.py > $OUT/band500.log 2>&1
timeout 200 python tools/small_band.py > $OUT/small_band.txt 2>&1; tail -8 $OUT/small_band.txt
