#!/bin/bash
# Round-2 iteration run: tests, smoke, default bench line, shard timings (contiguous / interleaved, with and without the
# exact-path launch), one ncu capture of video_kernel.   bash tools/gpu_iter.sh [tag]
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/timeline.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt; nproc >> $OUT/gpu.txt
stamp "pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -15 $OUT/pytest_gpu.txt
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
    print(o["config"]["workload"], "ms %.3f"%o.get("ms_per_step",0), "call %.3f"%(r.get("avg_launch_ms") or 0), "frac %.3f"%r.get("frac",0), "main %.3f"%(dk.get("avg_launch_ms") or 0), "slow", r.get("slow_path_pixels_per_launch"), "e2e %.3g"%(e.get("value") or 0), "verified", v.get("ok"), v.get("pixels_differing_from_oracle"), o.get("error"))
show(d)
for o in d.get("other_workloads",[]): show(o)
print("skipped", d.get("other_workloads_skipped"), "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
PY
stamp "shards"
timeout 300 python tools/small_band.py > $OUT/small_band.txt 2>&1; cat $OUT/small_band.txt
stamp "ncu full c5 video_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^video_kernel -s 1 -c 1 -f -o $OUT/ncu_video_c5 \
  python tools/prof_c5_ncu.py > $OUT/ncu_video_c5.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 40 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
head -45 $OUT/ncu_video_c5.summary.txt
stamp "done"
