#!/bin/bash
# Round-2 evidence run: tests, smoke, the default bench line (all workloads), the reference arm, the ncu launch list of
# the bench command and one `ncu --set full` capture per dominant kernel.  bash tools/gpu_r2i.sh [tag]
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/timeline.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt

stamp "pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu -s > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -4 $OUT/pytest_gpu.txt
stamp "smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; echo "exit $?" >> $OUT/smoke.txt
tail -2 $OUT/smoke.txt
stamp "bench default (c3 + other_workloads)"
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
grep -i "fail\|error\|Traceback" -A3 $OUT/bench_default.err | head -20
python - <<PY
import json
d=json.load(open("$OUT/bench_default.json"))
def show(o):
    r=o.get("roofline") or {}
    v=o.get("verified") or {}
    e=o.get("e2e") or {}
    dk=(r.get("dominant_kernel") or {})
    print(o["config"]["workload"], "ms %.3f"%o.get("ms_per_step",0), "call %.3f"%r.get("avg_launch_ms",0), "frac %.3f"%r.get("frac",0), "main %.3f"%dk.get("avg_launch_ms",0), "slow", r.get("slow_path_pixels_per_launch"), "e2e %.3g"%(e.get("value") or 0), "pageable", (e.get("pageable") or {}).get("h2d_gbs"), "cpu %.3g"%((o.get("cpu_baseline") or {}).get("value") or 0), "verified", v.get("ok"), v.get("pixels_differing_from_oracle"), o.get("error"))
show(d)
for o in d.get("other_workloads",[]): show(o)
print("skipped", d.get("other_workloads_skipped"), "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
PY
stamp "bench reference arm"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
cut -c1-300 $OUT/bench_reference.json
stamp "ncu launch list (bench c3)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_c3.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-others --no-verify > $OUT/launches_bench_c3.log 2>&1
stamp "ncu full c3 outlier_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_outlier_c3 \
  python tools/launch_times.py 4000 6000 200 0 > $OUT/ncu_outlier_c3.log 2>&1
stamp "ncu full a4 outlier_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_outlier_a4 \
  python tools/launch_times.py 4000 6000 200 0 4 > $OUT/ncu_outlier_a4.log 2>&1
stamp "ncu full c4 outlier_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^outlier_kernel -s 1 -c 1 -f -o $OUT/ncu_outlier_c4 \
  python tools/launch_times.py 2160 3840 1000 1 > $OUT/ncu_outlier_c4.log 2>&1
stamp "ncu launch lists a4 / c4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_a4.csv python tools/launch_times.py 4000 6000 200 0 4 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c4.csv python tools/launch_times.py 2160 3840 1000 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_a1.csv python tools/launch_times.py 2048 2048 200 0 3 > /dev/null 2>&1
stamp "ncu full a1 outlier_hard_kernel"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^outlier_hard_kernel -s 1 -c 1 -f -o $OUT/ncu_hard_a1 \
  python tools/launch_times.py 2048 2048 200 0 3 > $OUT/ncu_hard_a1.log 2>&1
stamp "band of 500 rows (one eighth of c3): launch list + per-band times"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_band500.csv python tools/small_band_once.py > $OUT/band500.log 2>&1
timeout 200 python tools/small_band.py > $OUT/small_band.txt 2>&1; tail -8 $OUT/small_band.txt
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
