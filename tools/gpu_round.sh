#!/bin/bash
# One gpurun call that refreshes every measured artefact: tests, smoke, bench lines, ncu launch list, ncu full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]    -> writes into gpurun_out/<tag>/
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/timeline.txt; }

stamp "pytest -m gpu"
timeout 600 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -3 $OUT/pytest_gpu.txt
stamp "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "exit $?" >> $OUT/smoke.txt
tail -2 $OUT/smoke.txt
stamp "bench c3 (default)"
timeout 400 python bench.py --steps 10 --warmup 3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err
cat $OUT/bench_c3.json | cut -c1-600
stamp "bench reference arm"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
for wl in c2-darker c2-lighter c4-outlier-rel-forward c1-minimal c5-video; do
  stamp "bench $wl"
  timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  cut -c1-300 $OUT/bench_$wl.json
done
stamp "quick_time c3 all modes"
timeout 300 python tools/quick_time.py 4000 6000 200 2 > $OUT/quick_c3.txt 2>&1
tail -6 $OUT/quick_c3.txt
stamp "ncu launch list (bench c3)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_c3.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench_c3.log 2>&1
stamp "ncu full c3 outlier_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_outlier_c3 \
  python tools/launch_times.py 4000 6000 200 0 > $OUT/ncu_outlier_c3.log 2>&1
stamp "ncu full c4 outlier_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_outlier_c4 \
  python tools/launch_times.py 2160 3840 1000 1 > $OUT/ncu_outlier_c4.log 2>&1
stamp "ncu full c4 outlier_hist_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:outlier_hist_kernel -s 1 -c 1 -f -o $OUT/ncu_hist_c4 \
  python tools/launch_times.py 2160 3840 1000 1 > $OUT/ncu_hist_c4.log 2>&1
stamp "ncu full c5 video_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^video_kernel -s 1 -c 1 -f -o $OUT/ncu_video_c5 \
  python tools/prof_c5_ncu.py > $OUT/ncu_video_c5.log 2>&1
stamp "summaries"
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 40 > $b.summary.txt 2>&1
  ncu -i $r --page raw --csv > $b.raw.csv 2>/dev/null
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  [ -n "$KEEP_REPS" ] || rm -f $r
done
stamp "done"
ls -la $OUT
