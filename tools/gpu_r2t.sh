#!/bin/bash
# chrono-video iteration: tests, c5 bench line, one ncu capture of video_kernel.   bash tools/gpu_r2t.sh [tag]
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/timeline.txt; }
stamp "pytest -m gpu"
timeout 1200 python -m pytest tests/ -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "exit $?" >> $OUT/pytest_gpu.txt
tail -6 $OUT/pytest_gpu.txt
stamp "bench c5"
timeout 600 python bench.py --workload c5-video --no-cpu > $OUT/bench_c5.json 2> $OUT/bench_c5.err
tail -3 $OUT/bench_c5.err
python - <<PY
import json
d=json.load(open("$OUT/bench_c5.json"))
print(d["config"]["workload"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "verified", d.get("verified"), "e2e", (d.get("e2e") or {}).get("value"))
PY
stamp "ncu full c5 video_kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^video_kernel -s 1 -c 1 -f -o $OUT/ncu_video_c5 \
  python tools/prof_c5_ncu.py > $OUT/ncu_video_c5.log 2>&1
for r in $OUT/*.ncu-rep; do
  b=${r%.ncu-rep}
  python tools/ncu_summary.py $r 30 > $b.summary.txt 2>&1
  ncu -i $r --page source --csv --print-source cuda,sass 2>/dev/null | gzip -9 > $b.source.csv.gz
  rm -f $r
done
head -34 $OUT/ncu_video_c5.summary.txt
stamp "done"
