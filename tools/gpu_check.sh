#!/bin/bash
# Quick confirmation after a kernel change: GPU parity tests, then timings of every mode at the c3 and c4 sizes and the c5 video.
TAG=${1:-chk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -15 $OUT/pytest_gpu.txt
timeout 300 python tools/quick_time.py 4000 6000 200 2 > $OUT/quick_c3.txt 2>&1; tail -6 $OUT/quick_c3.txt
timeout 300 python tools/quick_time.py 2160 3840 1000 2 > $OUT/quick_c4.txt 2>&1; tail -6 $OUT/quick_c4.txt
timeout 300 python bench.py --workload c5-video --steps 10 --warmup 3 --no-cpu > $OUT/bench_c5.json 2> $OUT/bench_c5.err; cut -c1-200 $OUT/bench_c5.json
if [ -n "$LAUNCHES" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c4.csv python tools/launch_times.py 2160 3840 1000 1 > $OUT/launches_c4.log 2>&1
grep -E "outlier" $OUT/launches_c4.csv | awk -F'","' '{print $5, $(NF)}' | tail -6
fi
