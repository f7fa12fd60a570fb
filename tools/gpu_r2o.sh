#!/bin/bash
# call slots (concurrency refactor), video direct kernel with deferred medians
OUT=gpurun_out/r2o
mkdir -p $OUT
export PYTHONUNBUFFERED=1
show() { python -c "
import sys,json
d=json.loads(open('$1').read().strip().splitlines()[-1]); r=d['roofline']; v=d.get('verified') or {}
print('$2', d['config']['workload'], 'ms %.3f'%d['ms_per_step'], 'call %.3f'%r.get('avg_launch_ms',0), 'frac %.3f'%r['frac'], 'main %.3f'%(r.get('dominant_kernel') or {}).get('avg_launch_ms',0), 'verified', v.get('ok'), v.get('pixels_differing_from_oracle'))
"; }
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -5 $OUT/pytest_gpu.txt
CHB_VIDEO_DIRECT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "video" > $OUT/pytest_video_direct.txt 2>&1; echo "exit $?" >> $OUT/pytest_video_direct.txt
tail -3 $OUT/pytest_video_direct.txt
CHB_VIDEO_DIRECT=0 timeout 400 python bench.py --workload c5-video --no-cpu --no-e2e --no-verify --steps 10 --warmup 3 > $OUT/slide_c5.json 2> $OUT/slide_c5.err; show $OUT/slide_c5.json sliding
CHB_VIDEO_DIRECT=1 timeout 400 python bench.py --workload c5-video --no-cpu --no-e2e --steps 10 --warmup 3 > $OUT/direct_c5.json 2> $OUT/direct_c5.err; show $OUT/direct_c5.json direct
CHB_VIDEO_DIRECT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^video_direct_kernel -s 1 -c 1 -f -o $OUT/ncu_video_c5 python tools/prof_c5_ncu.py > $OUT/ncu_video_c5.log 2>&1
CHB_VIDEO_DIRECT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_c5.csv python tools/prof_c5_ncu.py > /dev/null 2>&1
for wl in c3-outlier-abs-extreme c2-darker; do
  timeout 300 python bench.py --workload $wl --no-cpu --steps 10 --warmup 3 > $OUT/new_$wl.json 2> $OUT/new_$wl.err; show $OUT/new_$wl.json new
done
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 30 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
